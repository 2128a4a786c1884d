"""Latency of ONE loss+grad evaluation at the reference's own CPU-runnable size (BASELINE config #0:
examples/regression_1d.py, N = 100, D = 1, Linear + Rbf + Constant) and a few more small N: the B200 package
against the oracle port on this box's host cores.  At these sizes the evaluation is launch/host bound.

    python tools/bench_small.py [--sizes 100 400 1000]
"""
import argparse, json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def f(x):
    return np.sin(2.0 * np.pi * x) + np.cos(3.5 * np.pi * x) - 3.0 * x + 5.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", type=int, nargs="+", default=[100, 400, 1000])
    ap.add_argument("--reps", type=int, default=200)
    args = ap.parse_args()
    from oracle import gp_oracle as O
    from gptorch_b200 import kernels, _native as nv
    from gptorch_b200.models import GPR
    out = []
    for n in args.sizes:
        np.random.seed(42)
        x = np.linspace(0, 1, n).reshape((-1, 1))
        y = f(x) + 0.1 * np.random.randn(n, 1)
        model = GPR(x, y, kernels.Linear(1) + kernels.Rbf(1) + kernels.Constant(1))
        params = [p for p in model.parameters() if p.requires_grad]

        def step():
            for p in params:
                p.grad = None
            loss = model.loss()
            loss.backward()
            return loss

        for _ in range(20):
            step()
        torch.cuda.synchronize()
        nv.reset_launch_count()
        t0 = time.perf_counter()
        for _ in range(args.reps):
            loss = step()
        val = loss.item()          # one host read at the end; each step already syncs once on the potrf info
        torch.cuda.synchronize()
        gpu_ms = (time.perf_counter() - t0) / args.reps * 1e3
        launches = nv.launch_count() / args.reps
        # the optimiser bridge (what scipy's L-BFGS-B drives): parameters in, loss + flat gradient out
        theta = model._get_param_array()
        import contextlib, io
        from gptorch_b200 import settings
        bridge = {}
        with contextlib.redirect_stdout(io.StringIO()):
            for mode in (False, True):
                settings.cuda_graphs = mode
                for _ in range(5):
                    model._loss_and_grad(theta)
                t0 = time.perf_counter()
                for _ in range(args.reps):
                    model._loss_and_grad(theta)
                bridge[mode] = (time.perf_counter() - t0) / args.reps * 1e3
        bridge_ms, graphed_ms = bridge[False], bridge[True]
        graph_used = "_graph_eval" in model.__dict__
        # oracle port (torch CPU fp64) of the same composite model
        X, Y = torch.as_tensor(x), torch.as_tensor(y)
        raws = [torch.zeros(1, dtype=torch.float64, requires_grad=True) for _ in range(5)]

        def cpu_step():
            for r in raws:
                r.grad = None
            v_lin, ell, var, c, noise = [r.exp() for r in raws]
            K = O.cov_composite("k0 + k1 + k2", [("Linear", None, v_lin), ("Rbf", ell, var), ("Constant", None, c)], X)
            L = O.chol(K + noise * torch.eye(n, dtype=torch.float64))
            alpha = O.tri_solve(Y, L)
            l = 0.5 * alpha.pow(2).sum() + O.tri_logdet(L) + 0.5 * n * np.log(2 * np.pi)
            l.backward()
            return l

        for _ in range(5):
            cpu_step()
        t0 = time.perf_counter()
        reps = max(10, args.reps // 4)
        for _ in range(reps):
            cpu_step()
        cpu_ms = (time.perf_counter() - t0) / reps * 1e3
        out.append({"n": n, "gpu_ms_per_eval": gpu_ms, "optimizer_bridge_eager_ms_per_eval": bridge_ms, "optimizer_bridge_cuda_graph_ms_per_eval": graphed_ms, "graph_used": graph_used, "native_launches_per_eval": launches,
                    "cpu_oracle_ms_per_eval": cpu_ms, "cpu_threads": torch.get_num_threads(), "loss": val})
    print(json.dumps({"workload": "GPR Linear+Rbf+Constant D=1 loss+grad (BASELINE configs[0] family)", "results": out}))


if __name__ == "__main__":
    main()
