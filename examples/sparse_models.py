"""Sparse GP regression at scale on the B200 package: VFE (collapsed bound) and SVGP (minibatches), one GPU or several.

    python examples/sparse_models.py --model VFE  --num-points 1000000 --num-inducing 512
    python examples/sparse_models.py --model SVGP --num-points 4000000 --num-inducing 1024 --batch 32768 --host
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 examples/sparse_models.py --model VFE --num-points 2000000

The calls are the reference's (gptorch/models/sparse_gpr.py): construct the model from numpy arrays, a kernel and
inducing inputs, then loss() / backward() / predict_y().  What is specific to this package:
  * `model.distribute()` after `torch.distributed.init_process_group`: X / Y of each rank are ITS shard of the rows; the
    M x M statistics (VFE) or the gradients (SVGP, `dist.allreduce_grads`) are all-reduced over NCCL and every replicated
    parameter is broadcast from rank 0;
  * `SVGP(..., data_on_host=True)`: data sets beyond HBM stay in pinned host memory, minibatches are gathered and copied one
    step ahead;
  * `settings.vfe_phi_form` / `settings.svgp_quadratic_form` ("auto"): the cheaper order of operations when Kuu is well
    conditioned.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from gptorch_b200 import kernels, likelihoods  # noqa: E402
from gptorch_b200.dist import allreduce_grads, shard_rows  # noqa: E402
from gptorch_b200.models import SVGP, VFE  # noqa: E402


def make_data(n, d, seed):
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(d, 1, generator=torch.Generator().manual_seed(1), dtype=torch.float64)
    X = torch.rand(n, d, generator=g, dtype=torch.float64)
    Y = torch.sin(X @ w) + 0.1 * torch.randn(n, 1, generator=g, dtype=torch.float64)
    return X, Y, w


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="VFE", choices=["VFE", "SVGP"])
    ap.add_argument("--num-points", dest="n", type=int, default=1_000_000, help="rows over ALL ranks")
    ap.add_argument("--dim", dest="d", type=int, default=8)
    ap.add_argument("--num-inducing", dest="m", type=int, default=512)
    ap.add_argument("--batch", type=int, default=32768, help="SVGP minibatch per rank")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--lr", type=float, default=0.02)
    ap.add_argument("--host", action="store_true", help="SVGP: keep the data in pinned host memory")
    args = ap.parse_args()

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    lo, hi = shard_rows(args.n, rank, world)
    X, Y, w = make_data(hi - lo, args.d, seed=100 + rank)            # this rank's rows
    Z = torch.rand(args.m, args.d, generator=torch.Generator().manual_seed(2), dtype=torch.float64)
    kern = kernels.Matern52(args.d, ARD=True, length_scales=1.5 * np.ones(args.d))
    lik = likelihoods.Gaussian(variance=0.05)
    if args.model == "VFE":
        model = VFE(X.numpy(), Y.numpy(), kern, inducing_points=Z.numpy(), likelihood=lik)
    else:
        model = SVGP(X if args.host else X.numpy(), Y if args.host else Y.numpy(), kern, inducing_points=Z.numpy(),
                     likelihood=lik, batch_size=args.batch, data_on_host=args.host)
    if world > 1:
        model.distribute()

    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=args.lr)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for step in range(args.steps):
        opt.zero_grad(set_to_none=True)
        loss = model.loss()
        loss.backward()
        if world > 1 and args.model == "SVGP":
            allreduce_grads(model)                                   # VFE all-reduces inside its statistics node
        opt.step()
        if rank == 0 and (step % 5 == 0 or step == args.steps - 1):
            print("step %3d  loss %.6e" % (step, loss.item()), flush=True)
    torch.cuda.synchronize()
    sec = (time.perf_counter() - t0) / args.steps

    xs = np.random.RandomState(3).rand(1000, args.d)
    mean, var = model.predict_y(xs)                                   # numpy in, numpy out (as in the reference)
    truth = np.sin(xs @ w.numpy())
    if rank == 0:
        print("%s  N=%d (x%d ranks)  M=%d  %.1f ms per step   test RMSE %.4f   mean predictive std %.4f"
              % (args.model, args.n, world, args.m, sec * 1e3, float(np.sqrt(np.mean((mean - truth) ** 2))),
                 float(np.sqrt(var).mean())))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
