"""Distributed exact-GPR log-likelihood (block-column-cyclic Cholesky over the ranks), BASELINE config #5 shape.

    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_dist_gpr.py --n 131072
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_dist_gpr.py --n 8300 --panel 1024 --check
"""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--num-points", dest="n", type=int, default=131072)
    ap.add_argument("--dim", dest="d", type=int, default=8)
    ap.add_argument("--panel", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=0)
    ap.add_argument("--check", action="store_true", help="compare with the single-GPU GPR loss (N must fit one GPU)")
    ap.add_argument("--grad", action="store_true", help="time loss + backward (distributed inverse and gradient)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
    os.environ.setdefault("RANK", "0"); os.environ.setdefault("WORLD_SIZE", "1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    from bench import synth_regression
    from gptorch_b200 import kernels, likelihoods
    from gptorch_b200.models import GPR, DistributedGPR
    X, Y, _ = synth_regression(args.n, args.d)
    model = DistributedGPR(X.numpy(), Y.numpy(), kernels.Rbf(args.d, ARD=True), likelihood=likelihoods.Gaussian(variance=0.01),
                           panel=args.panel)
    params = [p for p in model.parameters() if p.requires_grad]
    def evaluate():
        if not args.grad:
            with torch.no_grad():
                return model.loss()
        for p in params: p.grad = None
        loss = model.loss()
        loss.sum().backward()
        return loss.detach()
    for _ in range(args.warmup): evaluate()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps): loss = evaluate()
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    out = {"metric": "distributed GPR loss+grad evals/s" if args.grad else "distributed GPR log-likelihood evals/s", "n": args.n, "d": args.d, "panel": args.panel, "n_gpus": world,
           "ms_per_eval": ms.item() / args.steps, "tflops_aggregate": args.n ** 3 / (1.0 if args.grad else 3.0) / (ms.item() / args.steps) / 1e9,
           "loss": loss.item(), "max_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
    if args.check and rank == 0:
        ref = GPR(X.numpy(), Y.numpy(), kernels.Rbf(args.d, ARD=True), likelihood=likelihoods.Gaussian(variance=0.01))
        if args.grad:
            l1 = ref.loss(); l1.sum().backward(); lref = l1.item()
            rp = [p for p in ref.parameters() if p.requires_grad]
            out["grad_check_rel"] = max(float((a.grad - b.grad).abs().max() / b.grad.abs().max()) for a, b in zip(params, rp))
            out["grads"] = [p.grad.flatten().tolist() for p in params]
        else:
            with torch.no_grad():
                lref = ref.loss().item()
        out["single_gpu_loss"] = lref
        out["check_rel"] = abs(loss.item() - lref) / abs(lref)
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()

if __name__ == "__main__":
    main()
