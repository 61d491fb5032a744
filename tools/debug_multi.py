import os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def worker(rank, world, initfile):
    import torch, torch.distributed as dist
    from riskaversetrajopt_b200 import _lib, dist as sd
    from riskaversetrajopt_b200.device_path import DevicePath
    from riskaversetrajopt_b200.drone import drone_params as dp
    from riskaversetrajopt_b200.drone.drone_utils import sample_uncertain_parameters
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"file://{initfile}", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    M = 1000
    np.random.seed(0)
    DWs, masses, obs_Qs = sample_uncertain_parameters('saa', M=M)
    us = np.random.RandomState(0).randn(20, 3)
    first, cnt = sd.shard_range(M, world, rank)
    for mode in sys.argv[1:] or ['sharded', 'peer', 'nccl']:
        p = DevicePath(_lib.SAA_DRONE, 'saa', 20, 0.1, cnt, M_global=M, sample_offset=first, device=rank)
        p.set_params_drone(dp, dp.OSQP_TOL)
        p.set_samples_drone(masses[first:first+cnt], DWs[first:first+cnt], obs_Qs[first:first+cnt])
        torch.cuda.synchronize(); print(rank, mode, "created", flush=True)
        asm = sd.ShardedAssembler(p, mode=mode)
        asm.bind_global_params(lambda q: q.set_params_drone(dp, dp.OSQP_TOL))
        torch.cuda.synchronize(); print(rank, mode, "assembler ok", asm.out if rank else None, flush=True)
        for it in (0, 2):
            b = asm.step(us, it)
            torch.cuda.synchronize(); dist.barrier()
            print(rank, mode, it, "step ok", flush=True)
    dist.destroy_process_group()

if __name__ == "__main__":
    import torch.multiprocessing as mp
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(worker, args=(2, os.path.join(d, "init")), nprocs=2, join=True)
