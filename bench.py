#!/usr/bin/env python
"""bench.py -- SAA linearize+assemble throughput (samples*steps/s), B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one SCP iteration's linearize+assemble over the whole sample set:
the drone problem of drone/drone_risk.py (S = 20, n_obs = 3) on a synthetic
sample set of M = 10^6 samples PER GPU (BASELINE.json config 4, weak scaling:
each rank owns its block of constraint rows; the only data-path collective is
the all-reduce of the 123 sample-mean sums).  Prints ONE JSON line.

  value   : M_total * S / t_step, inputs resident in HBM, device-timed (CUDA
            events on the launching stream, max over ranks).
  e2e     : same metric through the host-facing call (host `us` in, pinned host
            A.data u-block / u / l out, copies inside the timed region).
  roofline: algorithmic bytes (10 144 B per sample, SURVEY 8d / DESIGN.md) /
            kernel time vs the measured HBM peak in MEASURED_PEAKS.json.
  cpu_baseline: the oracle port on the host cores over a bounded sample.

--impl reference times the reference's own algorithm on the host CPU: Oracle-A
(vmapped forward-mode autodiff + dense packing + SciPy CSR->CSC, i.e.
drone/drone_risk.py:239-423 restated with torch.func because JAX is not in the
image) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

S = 20
BYTES_PER_SAMPLE = 10144          # 536 B read + 9608 B written
METRIC = "saa_linearize_assemble_throughput"
UNIT = "samples*steps/s"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _traffic():
    """dram bytes per launch of the assemble kernel from the committed ncu summary."""
    p = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("drone_assemble_kernel", {}).get("dram_bytes_per_launch_at_bench_M")
        except Exception:
            return None
    return None


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (NVML in a thread, 5 ms
    period; falls back to `nvidia-smi -lms`)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.sm, self.mask, self.max_mhz = index, [], 0, None
        self._stop = threading.Event()
        self.t = None
        self.mode = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons

            def loop():
                while not self._stop.is_set():
                    try:
                        self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        self.mask |= int(get_reasons(h))
                    except Exception:
                        pass
                    time.sleep(0.005)
            self.mode = "nvml"
            self.t = threading.Thread(target=loop, daemon=True)
            self.t.start()
        except Exception:
            self._start_smi()

    def _start_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.mode = "nvidia-smi"
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            inv = {v: k for k, v in self.REASONS.items()}

            def loop():
                for line in self.proc.stdout:
                    r = [x.strip() for x in line.split(",")]
                    try:
                        self.sm.append(float(r[0])); self.max_mhz = float(r[1])
                        for n, v in zip(names, r[2:6]):
                            if v.lower().startswith("active"):
                                self.mask |= inv[n]
                    except Exception:
                        pass
            self.t = threading.Thread(target=loop, daemon=True)
            self.t.start()
        except Exception:
            self.mode = None

    def mark(self):
        """number of samples so far (to delimit the timed region)"""
        return len(self.sm)

    def stop(self, lo=0, hi=None):
        self._stop.set()
        if self.mode == "nvidia-smi":
            self.proc.terminate()
        if self.t is not None:
            self.t.join(timeout=2)
        sm = self.sm[lo:hi] if self.sm[lo:hi] else self.sm
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(n for b, n in self.REASONS.items() if self.mask & b),
                "samples": len(sm), "source": self.mode,
                "window": "timed region" if self.sm[lo:hi] else "whole run"}


def synthetic_drone_samples(M, seed, device):
    """Config-4 samples, reference distributions (drone/drone_utils.py:61-93),
    generated on the device (synthetic data; the seed-exact NumPy stream is what the
    parity tests use)."""
    import torch
    from riskaversetrajopt_b200.drone import drone_params as dp
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    u = lambda *shape: torch.rand(*shape, generator=g, device=device, dtype=torch.float64)
    masses = dp.mass_nom - dp.mass_delta + 2 * dp.mass_delta * u(M)
    obs_Qs = torch.zeros((M, 3, 3, 3), dtype=torch.float64, device=device)
    radii = torch.as_tensor(dp.obs_radii, dtype=torch.float64, device=device)
    delta = -dp.obs_radii_deltas + 2 * dp.obs_radii_deltas * u(M, 3, 3)
    obs_Qs[:, :, [0, 1, 2], [0, 1, 2]] = 1.0 / (radii[None, :, None] + delta) ** 2
    DWs = float(np.sqrt(dp.dt)) * torch.randn((M, S, 6), generator=g, device=device, dtype=torch.float64)
    return DWs, masses, obs_Qs


def bench_us():
    from riskaversetrajopt_b200.drone import drone_params as dp
    us = np.zeros((S, 3))
    us[:, :2] = 0.01                                     # reference initial guess (:108-120)
    return us + 0.05 * np.random.RandomState(0).randn(S, 3)


# --------------------------------------------------------------------------- CPU legs
def cpu_port_baseline(M_s=None, budget_s=12.0):
    """Oracle port (analytic restatement) on the host cores over a bounded sample."""
    from oracle import cpu_port
    return cpu_port.time_drone(bench_us(), budget_s=budget_s, M_s=M_s)


def run_reference_arm(args):
    """The reference's own algorithm on the CPU (see module docstring)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle.oracle_a import DroneOracleA
    from riskaversetrajopt_b200.drone import drone_params as dp
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    M_s = args.ref_samples
    rs = np.random.RandomState(0)
    masses = rs.uniform(dp.mass_nom - dp.mass_delta, dp.mass_nom + dp.mass_delta, M_s)
    obs_Qs = np.zeros((M_s, 3, 3, 3))
    for o in range(3):
        for d in range(3):
            obs_Qs[:, o, d, d] = 1. / (dp.obs_radii[o] + rs.uniform(-dp.obs_radii_deltas, dp.obs_radii_deltas, M_s))**2
    DWs = np.sqrt(dp.dt) * rs.randn(M_s, S, 6)
    model = DroneOracleA(S, DWs, masses, obs_Qs, 'saa', 0.1)
    us = bench_us()
    for _ in range(args.warmup):
        model.get_constraints_coeffs(us, 2)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        A, l, u = model.get_constraints_coeffs(us, 2)
    dt_step = (time.perf_counter() - t0) / args.steps
    value = M_s * S / dt_step
    sample = (f"Oracle-A (torch.func vmap(jacfwd) + dense packing + SciPy CSR->CSC = the reference's "
              f"algorithm, drone/drone_risk.py:239-423) on {M_s} of the 10^6 samples per step; "
              f"dense matrix is O(M^2) so the reference cannot run the full workload")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "drone SAA linearize+assemble, S=20, n_obs=3 (BASELINE config 4)",
                   "samples_per_gpu": 1000000, "samples_per_step_on_cpu": M_s, "alpha": 0.1, "scp_iter": 2,
                   "method": "saa"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# --------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from riskaversetrajopt_b200 import _lib
    from riskaversetrajopt_b200.device_path import DevicePath
    from riskaversetrajopt_b200.drone import drone_params as dp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    M = args.samples_per_gpu
    M_global = M * world
    DWs, masses, obs_Qs = synthetic_drone_samples(M, seed=rank, device=device)
    path = DevicePath(_lib.SAA_DRONE, 'saa', S, 0.1, M, M_global=M_global, sample_offset=rank * M,
                      device=local_rank)
    path.set_params_drone(dp, dp.OSQP_TOL)
    path.set_samples_drone(masses, DWs, obs_Qs)
    # each rank keeps its own row block (compact matrix of its M samples) in its HBM
    path.set_output_geometry(M, 0)
    torch.cuda.synchronize()
    del DWs, masses, obs_Qs
    path._keep = []
    torch.cuda.empty_cache()
    us = bench_us()
    scp_iter = 2
    stream = torch.cuda.current_stream(device)

    def step():
        if world == 1:
            return path.assemble(us, scp_iter, finalize=True)
        b = path.assemble(us, scp_iter, finalize=False, write_shared=True)
        dist.all_reduce(path.mean_sums)                  # NCCL, 123 doubles
        path.finalize_means(b)
        return b

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    # ---- timed region: K steps, device events on the launching stream -------------
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    mark_lo = sampler.mark()
    ev[0].record(stream)
    for k in range(args.steps):
        step()
        ev[k + 1].record(stream)
    barrier()
    mark_hi = sampler.mark()
    total_ms = ev[0].elapsed_time(ev[-1])
    per_step = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    t = torch.tensor([total_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = M_global * S / (ms_per_step * 1e-3)

    # ---- kernel-only duration for the roofline (same stream, own events) ------------
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(args.steps)]
    torch.cuda.synchronize()
    for a, b_ in kev:
        a.record(stream)
        path.assemble(us, scp_iter, finalize=False)
        b_.record(stream)
    torch.cuda.synchronize()
    kernel_ms = float(np.mean([a.elapsed_time(b_) for a, b_ in kev]))
    clocks = sampler.stop(mark_lo, mark_hi) if rank == 0 else None

    # ---- e2e through the host-facing call: host us in, pinned host values out ------
    e2e = None
    try:
        n_rows, n_cols, nnz = path.pattern_sizes()
        n_var = 1140 * M + 177                           # u-column block of A.data (contiguous)
        b = path.buffers()
        hAx = torch.empty(n_var, dtype=torch.float64, pin_memory=True)
        hu = torch.empty(n_rows, dtype=torch.float64, pin_memory=True)
        hl = torch.empty(6, dtype=torch.float64, pin_memory=True)

        def e2e_step():
            step()
            hAx.copy_(b['Ax'][:n_var], non_blocking=True)
            hu.copy_(b['u'], non_blocking=True)
            hl.copy_(b['l'][:6], non_blocking=True)
            stream.synchronize()

        e2e_step()
        barrier()
        n_e2e = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        te = torch.tensor([(time.perf_counter() - t0) / n_e2e], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": M_global * S / float(te.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(us.nbytes),
               "d2h_bytes_per_step": int((n_var + n_rows + 6) * 8),
               "ms_per_step": float(te.item()) * 1e3, "steps": n_e2e,
               "what": "host us (480 B, kernel argument) -> pinned host A.data[u columns], u, l[:6]; "
                       "static y/slack/t columns and the CSC pattern stay on the host from setup"}
        del hAx, hu, hl
    except Exception as exc:  # e.g. not enough pinned host memory
        e2e = {"value": None, "unit": UNIT, "error": repr(exc)[:200]}

    # ---- tail-reduced subproblem, context beside the headline: what the host
    # QP solver needs when it is fed the K = 1.25 alpha M samples with the largest Z_i only
    tail = None
    if world == 1 and not args.no_tail:
        try:
            tail = measure_tail(path, us, scp_iter, stream, device)
        except Exception as exc:
            tail = {"error": repr(exc)[:200]}

    # ---- BASELINE target configuration at N > 1: M = 10^6 samples IN TOTAL, row blocks
    # delivered to rank 0 (fused peer-store gather over NVLink), reported beside the weak-scaling line
    target = None
    if world > 1 and not args.no_gather:
        try:
            target = measure_target_config(args, rank, world, local_rank, device)
        except Exception as exc:
            target = {"error": repr(exc)[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = _peaks()
    achieved = M * BYTES_PER_SAMPLE / (kernel_ms * 1e-3) / 1e9
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_port_baseline()
        except Exception as exc:
            cpu = {"value": None, "unit": UNIT, "error": repr(exc)[:200]}
    launches_per_step = 4       # drone_assemble + drone_axis_mean + reduce_partials + scatter_means
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "drone SAA linearize+assemble, S=20, n_obs=3 (BASELINE config 4)",
                   "samples_per_gpu": M, "samples_total": M_global, "alpha": 0.1, "scp_iter": scp_iter,
                   "method": "saa", "l2": "per-step output 9.6 GB/GPU >> 126 MB L2 (no flush needed)",
                   "row_blocks": "sharded (each rank keeps its block in HBM)" if world > 1 else "single GPU",
                   "collective": "all_reduce(123 f64) per step" if world > 1 else "none"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": _traffic(), "peak_source": peak_src,
                     "kernel": "drone_assemble_kernel<double,20,6>", "kernel_ms": kernel_ms,
                     "bytes_per_launch": M * BYTES_PER_SAMPLE,
                     "note": "event pair spans drone_assemble_kernel (97 % of it) plus drone_axis_mean_kernel "
                             "(z-axis mean rows, reads 168 of the 536 input bytes per sample) and the ~3 us "
                             "reduce_partials kernel; bytes = the step's algorithmic bytes"},
        "cpu_baseline": cpu,
        "e2e": e2e,
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clocks,
        "per_step_ms": [round(x, 4) for x in per_step],
    }
    if target is not None:
        out["target_config"] = target
    if tail is not None:
        out["tail_subproblem"] = tail
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def measure_tail(path, us, scp_iter, stream, device):
    """Per-iteration cost of the tail-reduced subproblem on one GPU: means + Z_i over all M samples,
    device selection of the K largest, gather, linearize+assemble of K samples; device-timed, and
    end to end with the reduced (A.data, l, u) copied to pinned host memory."""
    import torch
    from riskaversetrajopt_b200.tail import TailSubproblem
    t = TailSubproblem(path, margin=0.25)
    for _ in range(3):
        t.assemble(us, scp_iter)
    torch.cuda.synchronize()
    n = 10
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(n):
        t.assemble(us, scp_iter)
    b_.record(stream)
    torch.cuda.synchronize()
    dev_ms = a.elapsed_time(b_) / n
    t.get_constraints_coeffs(us, scp_iter, copy=False)
    t0 = time.perf_counter()
    for _ in range(5):
        A, l, u, idx = t.get_constraints_coeffs(us, scp_iter, copy=False)
    e2e_ms = (time.perf_counter() - t0) / 5 * 1e3
    return {"K": t.K, "M": path.M_local, "device_ms": dev_ms, "e2e_ms": e2e_ms,
            "d2h_bytes_per_step": int(A.data.nbytes + l.nbytes + u.nbytes + idx.nbytes),
            "note": "context, not the headline metric: QP restricted to the K samples with the largest "
                    "max-constraint value at the iterate (exact when the samples left out stay inactive); "
                    "e2e = host us -> (A.data, l, u, idx) of the reduced matrix in pinned host memory"}


def measure_target_config(args, rank, world, local_rank, device):
    """M = 10^6 samples in total over `world` GPUs (BASELINE.json target: < 10 ms per SCP
    iteration on 8 GPUs), for the three delivery modes of riskaversetrajopt_b200.dist."""
    import torch
    import torch.distributed as dist
    from riskaversetrajopt_b200 import _lib, dist as sd
    from riskaversetrajopt_b200.device_path import DevicePath
    from riskaversetrajopt_b200.drone import drone_params as dp
    torch.cuda.empty_cache()
    M_total = args.target_samples
    first, cnt = sd.shard_range(M_total, world, rank)
    us = bench_us()
    res = {"samples_total": M_total, "n_gpus": world, "unit": "ms per SCP iteration (linearize+assemble)"}
    for mode in ("sharded", "factored", "peer", "nccl"):
        if mode == "nccl" and M_total % world:
            continue
        DWs, masses, obs_Qs = synthetic_drone_samples(cnt, seed=100 + rank, device=device)
        path = DevicePath(_lib.SAA_DRONE, 'saa', S, 0.1, cnt, M_global=M_total, sample_offset=first,
                          device=local_rank)
        path.set_params_drone(dp, dp.OSQP_TOL)
        path.set_samples_drone(masses, DWs, obs_Qs)
        torch.cuda.synchronize()
        del DWs, masses, obs_Qs
        path._keep = []
        asm = sd.ShardedAssembler(path, mode=mode)
        asm.bind_global_params(lambda q: q.set_params_drone(dp, dp.OSQP_TOL))
        for _ in range(3):
            asm.step(us, 2)
        torch.cuda.synchronize(); dist.barrier()
        n = max(5, min(args.steps, 20))
        t0 = time.perf_counter()
        for _ in range(n):
            asm.step(us, 2)
        torch.cuda.synchronize(); dist.barrier()
        t = torch.tensor([(time.perf_counter() - t0) / n], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[mode + "_ms"] = float(t.item()) * 1e3
        if mode in ("peer", "factored"):
            asm.shared.close()
        del asm, path
        torch.cuda.empty_cache()
    # tail-reduced subproblem (stratified selection per rank) delivered to rank 0 over NVLink
    try:
        DWs, masses, obs_Qs = synthetic_drone_samples(cnt, seed=100 + rank, device=device)
        path = DevicePath(_lib.SAA_DRONE, 'saa', S, 0.1, cnt, M_global=M_total, sample_offset=first,
                          device=local_rank)
        path.set_params_drone(dp, dp.OSQP_TOL)
        path.set_samples_drone(masses, DWs, obs_Qs)
        torch.cuda.synchronize()
        del DWs, masses, obs_Qs
        path._keep = []
        asm = sd.ShardedTailAssembler(path, margin=0.25, mode='factored')
        for _ in range(3):
            asm.step(us, 2)
        torch.cuda.synchronize(); dist.barrier()
        n = max(5, min(args.steps, 20))
        t0 = time.perf_counter()
        for _ in range(n):
            asm.step(us, 2)
        torch.cuda.synchronize(); dist.barrier()
        t = torch.tensor([(time.perf_counter() - t0) / n], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res["tail_ms"] = float(t.item()) * 1e3
        res["tail_K_total"] = asm.K_total
        asm.close()
        del asm, path
        torch.cuda.empty_cache()
    except Exception as exc:
        res["tail_error"] = repr(exc)[:200]
    res["note"] = ("host-timed between barriers, max over ranks; 'sharded' leaves row blocks in their owners' HBM, "
                   "'peer' = kernels store into rank 0's arrays over NVLink (fused gather), "
                   "'factored' = kernels store the factored record (sensitivities + trajectory, 3.4 instead of 9.1 KB "
                   "per sample) into rank 0 over NVLink and rank 0 expands it to the CSC entries, "
                   "'nccl' = gather + merge kernel on rank 0, "
                   "'tail' = tail-reduced subproblem: every rank keeps the 1.25 alpha M_r samples of its shard with the "
                   "largest constraint values and stores their factored record into rank 0 over NVLink, rank 0 expands it into its K-sample matrix")
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--samples-per-gpu", type=int, default=1_000_000)
    ap.add_argument("--ref-samples", type=int, default=400)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gather", action="store_true")
    ap.add_argument("--no-tail", action="store_true")
    ap.add_argument("--target-samples", type=int, default=1_000_000)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
