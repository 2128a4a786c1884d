#!/usr/bin/env python
"""Run the UNMODIFIED reference (baseline/_ref, installed from /root/reference) on the GPU box: the host-CPU path
at BASELINE.json's named size and the vendor-library path (`model.cuda()` -> cuSOLVER/cuBLAS) on the B200.

    python tools/reference_box.py [--skip-cpu] [--out gpurun_out/r02_reference_box.json]

Writes one JSON file: per case the loss, every gradient, seconds per loss+grad and peak memory.  The N=32768 entries
are committed as tests/golden/gpr_n32768_reference.json (the parity pin of the headline config) and cited by bench.py.
None of this imports gptorch_b200.
"""
import argparse
import json
import os
import resource
import sys
import time
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
warnings.filterwarnings("ignore")

import gptorch  # noqa: E402  (the reference)
from gptorch import kernels as rk, likelihoods as rl  # noqa: E402
from gptorch.models import GPR, VFE, SVGP  # noqa: E402

assert os.path.realpath(gptorch.__file__).startswith(os.path.realpath(os.path.join(ROOT, "baseline", "_ref")))


def synth(n, d, seed=1234):
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(n, d, generator=g, dtype=torch.float64)
    w = torch.randn(d, 1, generator=g, dtype=torch.float64)
    Y = torch.sin(X @ w) + 0.1 * torch.randn(n, 1, generator=g, dtype=torch.float64)
    return X, Y, g


def grads_of(model):
    return {n: p.grad.detach().cpu().numpy().ravel().tolist() for n, p in model.named_parameters()
            if p.grad is not None and p.numel() <= 64}


def timed_eval(model, cuda, repeats, *args):
    """(loss, grads, best seconds, all seconds) of model.loss(*args) + backward()."""
    secs = []
    for _ in range(repeats):
        for p in model.parameters():
            p.grad = None
        if cuda:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        loss = model.loss(*args)
        loss.backward()
        if cuda:
            torch.cuda.synchronize()
        secs.append(time.perf_counter() - t0)
    return loss, grads_of(model), min(secs), secs


def gpr_model(kind, n, d, ell=None, var=1.0, noise=0.01):
    X, Y, _ = synth(n, d)
    kern = getattr(rk, kind)(d, ARD=True, length_scales=None if ell is None else np.array(ell, dtype=np.float64), variance=var)
    return GPR(X.numpy(), Y.numpy(), kern, likelihood=rl.Gaussian(variance=noise))


def run_gpr(kind, n, d, cuda, repeats, **hyper):
    model = gpr_model(kind, n, d, **hyper)
    if cuda:
        model.cuda()
        torch.cuda.reset_peak_memory_stats()
    loss, gr, best, secs = timed_eval(model, cuda, repeats)
    out = {"kind": kind, "n": n, "d": d, "device": "cuda" if cuda else "cpu", "loss": float(loss.item()), "grads": gr,
           "seconds_best": best, "seconds_all": secs}
    if cuda:
        out["peak_gb"] = torch.cuda.max_memory_allocated() / 1e9
        del model, loss
        torch.cuda.empty_cache()
    else:
        out["threads"] = torch.get_num_threads()
        out["maxrss_gb"] = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e6
    return out


def run_vfe(n, d, m, cuda, repeats):
    X, Y, g = synth(n, d)
    Z = X[torch.randperm(n, generator=g)[:m]]
    model = VFE(X.numpy(), Y.numpy(), rk.Rbf(d, ARD=True), inducing_points=Z.numpy(), likelihood=rl.Gaussian(variance=0.01))
    if cuda:
        model.cuda()
        torch.cuda.reset_peak_memory_stats()
    loss, gr, best, secs = timed_eval(model, cuda, repeats)
    out = {"n": n, "d": d, "m": m, "device": "cuda" if cuda else "cpu", "loss": float(loss.item()), "grads": gr,
           "seconds_best": best, "seconds_all": secs}
    if cuda:
        out["peak_gb"] = torch.cuda.max_memory_allocated() / 1e9
    return out


def run_svgp(n, d, m, batch, cuda, repeats):
    X, Y, g = synth(n, d)
    Z = X[torch.randperm(n, generator=g)[:m]]
    idx = torch.randperm(n, generator=g)[:batch]
    np.random.seed(0)
    model = SVGP(X.numpy(), Y.numpy(), rk.Matern52(d, ARD=True), inducing_points=Z.numpy(),
                 likelihood=rl.Gaussian(variance=0.01), batch_size=batch)
    xb, yb = X[idx], Y[idx]
    if cuda:
        model.cuda()
        xb, yb = xb.cuda(), yb.cuda()
        torch.cuda.reset_peak_memory_stats()
    loss, gr, best, secs = timed_eval(model, cuda, repeats, xb, yb)
    out = {"n": n, "d": d, "m": m, "batch": batch, "device": "cuda" if cuda else "cpu", "loss": float(loss.item()),
           "grads": gr, "seconds_best": best, "seconds_all": secs}
    if cuda:
        out["peak_gb"] = torch.cuda.max_memory_allocated() / 1e9
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_reference_box.json"))
    ap.add_argument("--n", type=int, default=32768)
    args = ap.parse_args()
    import psutil
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    res = {"host": {"cores": cores, "ram_gb": psutil.virtual_memory().total / 1e9,
                    "ram_available_gb": psutil.virtual_memory().available / 1e9},
           "torch": torch.__version__, "gpu": torch.cuda.get_device_name(0) if torch.cuda.is_available() else None,
           "reference": os.path.dirname(gptorch.__file__)}

    def save():
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)

    cuda = torch.cuda.is_available()
    # --- reference vs itself (CPU/MKL against CUDA/cuSOLVER) on small cases: the reference's own noise floor ----
    self_cases = []
    for kind in ("Rbf", "Exp", "Matern32", "Matern52"):
        for n in (1024, 4096):
            c = run_gpr(kind, n, 8, False, 1)
            entry = {"kind": kind, "n": n, "cpu": c}
            if cuda:
                entry["cuda"] = run_gpr(kind, n, 8, True, 1)
            self_cases.append(entry)
            print(kind, n, "cpu", repr(c["loss"]), "cuda", repr(entry.get("cuda", {}).get("loss")), flush=True)
    res["self_consistency"] = self_cases
    save()
    # --- vendor-library path on the B200 -----------------------------------------------------------------------
    if cuda:
        run_gpr("Rbf", 2048, 8, True, 1)   # warm up cuSOLVER/cuBLAS handles
        for n in (8192, 16384, args.n):
            r = run_gpr("Rbf", n, 8, True, 3)
            res["gpr_cuda_n%d" % n] = r
            print("reference GPR cuda N=%d loss %r best %.3f s peak %.1f GB" % (n, r["loss"], r["seconds_best"], r["peak_gb"]), flush=True)
            save()
        try:
            r = run_vfe(100000, 16, 1024, True, 3)
            res["vfe_cuda_n100000"] = r
            print("reference VFE cuda N=1e5 loss %r best %.3f s" % (r["loss"], r["seconds_best"]), flush=True)
            r = run_vfe(1250000, 16, 1024, True, 2)
            res["vfe_cuda_n1250000"] = r
            print("reference VFE cuda N=1.25e6 loss %r best %.3f s peak %.1f GB" % (r["loss"], r["seconds_best"], r["peak_gb"]), flush=True)
        except Exception as e:  # noqa: BLE001
            res["vfe_cuda_error"] = repr(e)
            print("VFE cuda failed:", e, flush=True)
        torch.cuda.empty_cache()
        save()
        try:
            r = run_svgp(262144, 32, 2048, 65536, True, 3)
            res["svgp_cuda_b65536"] = r
            print("reference SVGP cuda B=65536 loss %r best %.3f s peak %.1f GB" % (r["loss"], r["seconds_best"], r["peak_gb"]), flush=True)
        except Exception as e:  # noqa: BLE001
            res["svgp_cuda_error"] = repr(e)
            print("SVGP cuda failed:", e, flush=True)
        torch.cuda.empty_cache()
        save()
    # --- host-CPU path at the named size (needs ~62-77 GB of RAM and minutes) ---------------------------------------
    if not args.skip_cpu:
        avail = psutil.virtual_memory().available / 1e9
        need = 9.5 * 8 * args.n * args.n / 1e9
        if avail >= need:
            r = run_gpr("Rbf", args.n, 8, False, 1)
            res["gpr_cpu_n%d" % args.n] = r
            print("reference GPR cpu N=%d loss %r %.1f s maxrss %.1f GB (%d threads)"
                  % (args.n, r["loss"], r["seconds_best"], r["maxrss_gb"], r["threads"]), flush=True)
        else:
            res["gpr_cpu_n%d" % args.n] = {"skipped": "host has %.0f GB available, the reference needs ~%.0f GB" % (avail, need)}
            print(res["gpr_cpu_n%d" % args.n], flush=True)
        save()
    print("wrote", args.out)


if __name__ == "__main__":
    main()
