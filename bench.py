#!/usr/bin/env python
"""Headline benchmark: fp64 GPR loss+grad evaluations per second at N=32768, D=8 on B200 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--num-points 32768]

One "step" = one full evaluation of model.loss() + loss.backward() for GPR with an Rbf-ARD kernel on the
synthetic regression problem of BASELINE.md section 3 (covariance build, Cholesky, solves, log-det, and the
analytic gradient through (L L^T)^-1).  Prints ONE JSON line (rank 0).

* value      : evaluations/s with X, Y resident in HBM, CUDA-event timed, max over ranks.
* e2e        : the same through the public API with HOST inputs: per step X, Y are copied from pinned host
               memory into the model, loss()+backward() run, loss and gradients are read back to the host.
* roofline   : the O(N^3) part (gpb potrf + potri, > 95 % DMMA GEMM kernel) against the FP64 peak measured
               live with a cuBLAS DGEMM; MEASURED_PEAKS.json carries no FP64 figure.
* cpu_baseline / --impl reference : the oracle port of the reference's CPU path (torch CPU / MKL, all host
               threads) on a bounded sample (N=4096 or 8192), scaled by N^3 to the named shape.
* --gpus N>1 : the GPR evaluation does not shard at this size (DESIGN.md "replicas only"): every
               rank runs an independent replica; value = N * K / max-rank time.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fp64 GPR loss+grad evals/sec @N=32768,D=8; Cholesky FP64 TFLOP/s vs peak"
D_IN = 8
FP64_DMMA_PROBE_TFLOPS = 37.0   # tools/probe_fp64.cu on this pool's B200 (profiles/r01_fp64_peak.txt)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--num-points", dest="n", type=int, default=32768, help="training points (the named config is 32768)")
    ap.add_argument("--cpu-sample-n", type=int, default=0, help="oracle sample size (0 = choose by core count)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------
# CPU side: the oracle port of the reference path, bounded sample
# ------------------------------------------------------------------------------------------------------
def cpu_eval_seconds(n, repeats=1):
    from oracle import gp_oracle as O
    X, Y, _ = O.synth_regression(n, D_IN)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        O.gpr_loss_and_grads("Rbf", X, Y, np.ones(D_IN), 1.0, 0.01)
        best = min(best, time.perf_counter() - t0)
    return best


def default_cpu_sample(cores):
    """Sample size of the CPU leg: about 10-30 s of host work per evaluation (the reference path costs
    ~4.3 N^3 flop and ~9 N^2 fp64 temporaries), bounded by the host's free memory."""
    n = 12288 if cores >= 16 else (8192 if cores >= 8 else 4096)
    try:
        import psutil
        free = psutil.virtual_memory().available
        while n > 4096 and 16 * 8 * n * n > free:
            n -= 4096
    except Exception:
        n = min(n, 8192)
    return n


def cpu_baseline(n_full, sample_n=0):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if not sample_n:
        sample_n = default_cpu_sample(cores)
    sample_n = min(sample_n, n_full)
    cpu_eval_seconds(1024)  # warm up MKL / autograd
    sec = cpu_eval_seconds(sample_n)
    scale = (n_full / sample_n) ** 3
    return {
        "value": 1.0 / (sec * scale),
        "unit": "evals/s",
        "cores": torch.get_num_threads(),
        "kind": "port",
        "sample": "oracle/gp_oracle.py (torch CPU fp64, MKL) loss+grad at N=%d, D=%d: %.2f s measured; scaled by "
                  "(N/%d)^3 = %.0fx to N=%d" % (sample_n, D_IN, sec, sample_n, scale, n_full),
        "sample_seconds": sec,
    }


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample_n = args.cpu_sample_n or default_cpu_sample(cores)
    sample_n = min(sample_n, args.n)
    for _ in range(max(args.warmup, 1)):
        cpu_eval_seconds(1024)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_eval_seconds(sample_n)
    sec = (time.perf_counter() - t0) / args.steps
    scale = (args.n / sample_n) ** 3
    value = 1.0 / (sec * scale)
    sample = ("oracle port of the reference CPU path (torch CPU fp64, MKL, %d threads): loss+grad at N=%d timed "
              "%.2f s/step, scaled by (N/%d)^3 = %.0fx to N=%d" % (torch.get_num_threads(), sample_n, sec, sample_n, scale, args.n))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / value, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "GPR Rbf-ARD fp64 N=%d D=%d dy=1 loss+grad (configs[1])" % (args.n, D_IN)},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax = float(f[2])
            except ValueError:
                continue
            for name, flag in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        busy = [v for v in sm if v > 0.5 * (smax or 1)] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def fp64_peak_live():
    """Measured FP64 peak on this device: best of 5 cuBLAS DGEMMs (the same way MEASURED_PEAKS.json is made)."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    torch.cuda.empty_cache()
    return 2.0 * n ** 3 / best / 1e9


PROBE_M, PROBE_K = 16384, 2048
PROBE_TRAFFIC_BYTES = 5.50e9   # dram__bytes_read.sum + dram__bytes_write.sum, profiles/r01_ncu_syrk_16384x2048.txt


def dominant_launch_probe():
    """One launch shape of the dominant kernel timed on its own with CUDA events: the SYRK trailing update
    (lower tiles) m = n = 16384, k = 2048 -- the shape profiled with ncu --set full in profiles/."""
    from gptorch_b200 import _native as nv
    A = torch.randn(PROBE_M, PROBE_K, dtype=torch.float64, device="cuda")
    C = torch.randn(PROBE_M, PROBE_M, dtype=torch.float64, device="cuda")
    run = lambda: nv.gemm(nv.GEMM_NT, A, A, alpha=-1.0, beta=1.0, C=C, lower_only=True)  # noqa: E731
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    times = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = sum(times) / len(times)
    tiles = (PROBE_M // 128) * (PROBE_M // 128 + 1) // 2
    flop = 2.0 * tiles * 128 * 128 * PROBE_K
    del A, C
    torch.cuda.empty_cache()
    return {"shape": "syrk lower m=n=%d k=%d (%d tiles of 128x128; C is 2.1 GB > L2)" % (PROBE_M, PROBE_K, tiles),
            "ms": ms, "flop": flop, "tflops": flop / ms / 1e9}


def synth_regression(n, d, seed=1234):
    """SURVEY 8d's synthetic inputs (CPU generator, so the reference's loss pins apply): X ~ U[0,1)^d,
    Y = sin(X w) + 0.1 eps.  Stated here rather than imported: oracle/ is only the checker / CPU baseline."""
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(n, d, generator=g, dtype=torch.float64)
    w = torch.randn(d, 1, generator=g, dtype=torch.float64)
    Y = torch.sin(X @ w) + 0.1 * torch.randn(n, 1, generator=g, dtype=torch.float64)
    return X, Y, g


def build_model(n, device):
    from gptorch_b200 import kernels, likelihoods
    from gptorch_b200.models import GPR
    X, Y, _ = synth_regression(n, D_IN)
    model = GPR(X.numpy(), Y.numpy(), kernels.Rbf(D_IN, ARD=True), likelihood=likelihoods.Gaussian(variance=0.01))
    return model, X, Y


def one_eval(model):
    for p in model.parameters():
        p.grad = None
    loss = model.loss()
    loss.backward()
    return loss


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from gptorch_b200 import _native as nv
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    n = args.n
    model, X, Y = build_model(n, device)
    params = [p for p in model.parameters() if p.requires_grad]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing -----------------------------------------------------------
    for _ in range(args.warmup):
        loss = one_eval(model)
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    timer = nv.PhaseTimer()
    nv.install_timer(timer)
    nv.reset_launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = one_eval(model)
    e1.record()
    barrier()
    launches = nv.launch_count()
    nv.install_timer(None)
    ms = e0.elapsed_time(e1)
    phases = timer.totals_ms()
    clocks = sampler.stop() if rank == 0 else None
    loss_value = float(loss.item())

    # ---------------- end-to-end through the public API with host inputs ---------------------------------
    x_host = X.clone().pin_memory()
    y_host = Y.clone().pin_memory()
    e2e_steps = max(1, min(args.steps, 3))
    one_eval(model)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        model.X.copy_(x_host, non_blocking=True)
        model.Y.copy_(y_host, non_blocking=True)
        loss = one_eval(model)
        host_loss = loss.detach().cpu()
        host_grads = [p.grad.detach().cpu() for p in params]
    torch.cuda.synchronize()
    e2e_sec = time.perf_counter() - t0
    h2d = (x_host.numel() + y_host.numel()) * 8
    d2h = (host_loss.numel() + sum(g.numel() for g in host_grads)) * 8

    t = torch.tensor([ms, e2e_sec * 1000.0], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max = t[0].item(), t[1].item()
    if rank != 0:
        return

    steps = args.steps
    value = world * steps / (ms_max / 1000.0)
    e2e_value = world * e2e_steps / (e2e_ms_max / 1000.0)
    peak = fp64_peak_live()
    potrf_ms = phases.get("potrf", 0.0) / steps
    potri_ms = phases.get("potri", 0.0) / steps
    n3 = float(n) ** 3
    chol_tflops = n3 / 3.0 / potrf_ms / 1e9 if potrf_ms else None
    o3_tflops = n3 / (potrf_ms + potri_ms) / 1e9 if (potrf_ms + potri_ms) else None
    probe = dominant_launch_probe()
    # pins from the reference for the sizes the oracle could run
    pins = {1024: -606.3903292756472, 2048: -1420.2752146205817, 4096: -2680.7933915936683,
            8192: -6511.334472842767, 16384: -13224.865836863326}
    line = {
        "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world, "steps": steps, "warmup": args.warmup,
        "ms_per_step": ms_max / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "GPR Rbf-ARD fp64 N=%d D=%d dy=1 loss+grad (configs[1])" % (n, D_IN),
                   "parallelism": "replicas only (x%d)" % world if world > 1 else "single GPU",
                   "l2": "inputs_exceed_l2 (the N^2 covariance/factor buffer is %.1f GB)" % (8.0 * n * n / 1e9),
                   "loss": loss_value, "loss_pin": pins.get(n)},
        "chol_tflops": chol_tflops,
        "phases_ms_per_step": {k: v / steps for k, v in sorted(phases.items())},
        "roofline": {"bound": "tensor", "achieved": o3_tflops, "peak": peak, "unit": "TFLOP/s",
                     "frac": (o3_tflops / peak) if o3_tflops else None,
                     "traffic": PROBE_TRAFFIC_BYTES,
                     "kernel": "gemm_dmma_kernel (FP64 DMMA.8x8x4 fed by TMA) -- > 95 % of gpb_potrf_lower + gpb_potri_lower",
                     "algorithmic": "achieved = N^3 flop per eval (potrf N^3/3 + potri 2N^3/3) / CUDA-event time of those two "
                                    "phases inside the timed steps (all their launches, including the latency-bound ones)",
                     "peak_source": "measured live: cuBLAS DGEMM 8192^3 best of 5 (MEASURED_PEAKS.json has no FP64 entry); "
                                    "DMMA issue-rate probe on this pool: %.1f TFLOP/s" % FP64_DMMA_PROBE_TFLOPS,
                     "chol_frac": (chol_tflops / peak) if chol_tflops else None,
                     "launch_probe": probe,
                     "traffic_note": "dram__bytes_read+write of ONE launch of the probe shape from ncu --set full "
                                     "(profiles/r01_ncu_syrk_16384x2048.txt); algorithmic bytes of that launch: 2.43e9"},
        "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if not args.no_cpu_baseline and world == 1:     # reported on rank 0 of the single-GPU run only
        line["cpu_baseline"] = cpu_baseline(n, args.cpu_sample_n)
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (gptorch_b200 has no CPU path); use --impl reference for the CPU arm")
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
