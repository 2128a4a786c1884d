"""VFE loss+grad throughput (BASELINE config #3 shape: D=16, M=1024, Rbf-ARD), rows sharded over the ranks.

    python tools/bench_vfe.py --n-per-gpu 1250000            # one GPU's shard of N = 1e7
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_vfe.py --n-per-gpu 1250000
Prints one JSON line on rank 0 (evals/s over the GLOBAL N = n_per_gpu * world; max-over-ranks CUDA-event time).
"""
import argparse, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-per-gpu", type=int, default=1250000)
    ap.add_argument("--num-inducing", dest="m", type=int, default=1024)
    ap.add_argument("--dim", dest="d", type=int, default=16)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--check", action="store_true", help="compare the sharded loss/grads with a single-process run (small N)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    from gptorch_b200 import kernels, likelihoods, _native as nv
    from gptorch_b200.models import VFE
    from gptorch_b200.dist import shard_rows
    n_global = args.n_per_gpu * world
    g = torch.Generator().manual_seed(1234)
    # inducing points and the regression weights are global; every rank generates its own rows
    w = torch.randn(args.d, 1, generator=g, dtype=torch.float64)
    Z = torch.rand(args.m, args.d, generator=g, dtype=torch.float64)
    gr = torch.Generator().manual_seed(1000 + rank)
    X = torch.rand(args.n_per_gpu, args.d, generator=gr, dtype=torch.float64)
    Y = torch.sin(X @ w) + 0.1 * torch.randn(args.n_per_gpu, 1, generator=gr, dtype=torch.float64)
    model = VFE(X.numpy(), Y.numpy(), kernels.Rbf(args.d, ARD=True), inducing_points=Z.numpy(),
                likelihood=likelihoods.Gaussian(variance=0.01))
    if world > 1:
        model.distribute()
    def step():
        for p in model.parameters(): p.grad = None
        loss = model.loss(); loss.backward(); return loss
    for _ in range(args.warmup): loss = step()
    timer = nv.PhaseTimer(); nv.install_timer(timer); nv.reset_launch_count()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps): loss = step()
    e1.record()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    phases = timer.totals_ms(); nv.install_timer(None)
    if rank == 0:
        sec = ms.item() / 1000 / args.steps
        flop = 5.0 * n_global * args.m ** 2 + 6.0 * n_global * args.m * args.d
        print(json.dumps({"metric": "VFE loss+grad evals/s", "n_global": n_global, "m": args.m, "d": args.d, "n_gpus": world,
                          "value": 1.0 / sec, "ms_per_eval": sec * 1000, "rows_per_s": n_global / sec,
                          "algorithmic_tflops": flop / sec / 1e12, "loss": loss.item(),
                          "grad_ell0": model.kernel.length_scales.grad[0].item(), "grad_Z00": model.Z.grad[0, 0].item(),
                          "phases_ms_per_eval": {k: v / args.steps for k, v in sorted(phases.items())},
                          "launches_per_eval": nv.launch_count() / args.steps}), flush=True)
    if args.check:
        # single-process evaluation on ALL rows (rank 0 regenerates every rank's shard) vs the sharded result
        grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
        if rank == 0:
            Xs, Ys = [], []
            for r in range(world):
                g2 = torch.Generator().manual_seed(1000 + r)
                Xr = torch.rand(args.n_per_gpu, args.d, generator=g2, dtype=torch.float64)
                Yr = torch.sin(Xr @ w) + 0.1 * torch.randn(args.n_per_gpu, 1, generator=g2, dtype=torch.float64)
                Xs.append(Xr); Ys.append(Yr)
            full = VFE(torch.cat(Xs).numpy(), torch.cat(Ys).numpy(), kernels.Rbf(args.d, ARD=True), inducing_points=Z.numpy(),
                       likelihood=likelihoods.Gaussian(variance=0.01))
            lf = full.loss(); lf.backward()
            rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
            out = {"check_loss_rel": abs(loss.item() - lf.item()) / abs(lf.item())}
            for n, p in full.named_parameters():
                if p.grad is not None: out["check_grad_rel/" + n] = rel(grads[n], p.grad)
            print(json.dumps(out), flush=True)
    if world > 1: dist.destroy_process_group()

if __name__ == "__main__":
    main()
