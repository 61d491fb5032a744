for cfg in "50 1.6 0.1" "25 1.6 0.1" "100 1.6 0.1" "50 1.8 0.1" "50 1.6 0.01" "20 1.7 0.05"; do set -- $cfg
python examples/large_scp_device.py --M 100000 --iters 8 --eps 1e-4 --rho-interval $1 --relax $2 --rho0 $3 2>/dev/null | grep "^{" | python -c "
import json,sys; d=json.loads(sys.stdin.read()); its=[i['admm_iters'] for i in d['iterations']]; print('$cfg', its, sum(its), round(sum(i['solve_ms'] for i in d['iterations'])), d['final']['satisfied_fraction'], round(d['iterations'][-1]['left_out_margin'],3), round(d['iterations'][-1]['L2_error'],3))"
done
