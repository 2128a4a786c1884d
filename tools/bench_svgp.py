"""SVGP minibatch loss+grad throughput (BASELINE config #4 shape: Matern52-ARD, D=32, M=2048, batch 65536),
data-parallel over the ranks: every rank evaluates its own minibatch, gradients are all-reduced as one flat buffer.

    python tools/bench_svgp.py --batch 65536
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_svgp.py --batch 65536
"""
import argparse, json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-per-gpu", type=int, default=262144)
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--num-inducing", dest="m", type=int, default=2048)
    ap.add_argument("--dim", dest="d", type=int, default=32)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--host", action="store_true",
                    help="keep X, Y in pinned HOST memory and stream minibatches (SVGP(data_on_host=True)); with "
                         "--n-per-gpu 100000000 this is BASELINE config #4 at its full N = 1e8 (25.6 GB of inputs)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    from gptorch_b200 import kernels, likelihoods, _native as nv
    from gptorch_b200.models import SVGP
    from gptorch_b200.dist import allreduce_grads
    g = torch.Generator().manual_seed(1234)
    w = torch.randn(args.d, 1, generator=g, dtype=torch.float64)
    Z = torch.rand(args.m, args.d, generator=g, dtype=torch.float64)
    gr = torch.Generator().manual_seed(1000 + rank)
    import time
    t_gen = time.perf_counter()
    X = torch.empty(args.n_per_gpu, args.d, dtype=torch.float64)
    Y = torch.empty(args.n_per_gpu, 1, dtype=torch.float64)
    step_rows = 1 << 22                     # generate in slabs: bounded temporaries at N = 1e8
    for s0 in range(0, args.n_per_gpu, step_rows):
        xs = X[s0: s0 + step_rows]
        xs.copy_(torch.rand(xs.shape[0], args.d, generator=gr, dtype=torch.float64))
        Y[s0: s0 + step_rows] = torch.sin(xs @ w) + 0.1 * torch.randn(xs.shape[0], 1, generator=gr, dtype=torch.float64)
    t_gen = time.perf_counter() - t_gen
    np.random.seed(rank)
    model = SVGP(X if args.host else X.numpy(), Y if args.host else Y.numpy(),
                 kernels.Matern52(args.d, ARD=True, length_scales=2.0 * np.ones(args.d)), inducing_points=Z.numpy(),
                 likelihood=likelihoods.Gaussian(variance=0.01), batch_size=args.batch, data_on_host=args.host)
    if world > 1:
        model.distribute()
    def step():
        for p in model.parameters(): p.grad = None
        loss = model.loss(); loss.backward()
        if world > 1: allreduce_grads(model)
        return loss
    for _ in range(args.warmup): loss = step()
    nv.reset_launch_count()
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
    if rank == 0:
        sec = ms.item() / 1000 / args.steps
        print(json.dumps({"metric": "SVGP minibatch loss+grad steps/s", "batch_per_gpu": args.batch, "m": args.m, "d": args.d,
                          "n_gpus": world, "value": 1.0 / sec, "ms_per_step": sec * 1000, "points_per_s": world * args.batch / sec,
                          "n_per_gpu": args.n_per_gpu, "data": "pinned host memory, minibatches streamed (HostBatchStream)" if args.host else "resident in HBM",
                          "host_gb": (X.numel() + Y.numel()) * 8 / 1e9, "generate_s": t_gen,
                          "loss": loss.item(), "launches_per_step": nv.launch_count() / args.steps,
                          "max_mem_gb": torch.cuda.max_memory_allocated() / 1e9}), flush=True)
    if world > 1: dist.destroy_process_group()

if __name__ == "__main__":
    main()
