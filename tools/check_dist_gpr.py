"""Size-independent correctness checks of the distributed exact GPR at BASELINE config #5 (N = 131072 on 8 GPUs), where no
single-GPU or reference value exists:

  (1) inverse identity  Ky^-1 (Ky v) = v  for a random v, with Ky v formed slab by slab from the covariance kernel and
      Ky^-1 applied from the distributed inverse (lower block triangle, mirrored);
  (2) gradient vs a central finite difference of the loss along a random direction of the raw hyper-parameters;
  (3) the loss against the value recorded in round 1 (profiles/r01_multigpu.json).

    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 tools/check_dist_gpr.py --num-points 131072

The matvecs of the checker use torch (it is test infrastructure, not the product path).  Prints one JSON line on rank 0.
"""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--num-points", dest="n", type=int, default=131072)
    ap.add_argument("--panel", type=int, default=1024)
    ap.add_argument("--eps", type=float, default=1e-4)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29534")
    os.environ.setdefault("RANK", "0"); os.environ.setdefault("WORLD_SIZE", "1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    from bench import synth_regression
    from gptorch_b200 import kernels, likelihoods, _native as nv
    from gptorch_b200.models import DistributedGPR
    n, w = args.n, args.panel
    X, Y, _ = synth_regression(n, 8)
    model = DistributedGPR(X.numpy(), Y.numpy(), kernels.Rbf(8, ARD=True), likelihood=likelihoods.Gaussian(variance=0.01), panel=w)
    params = [p for p in model.parameters() if p.requires_grad]
    out = {"n": n, "n_gpus": world, "panel": w}

    # ---- (1) inverse identity ---------------------------------------------------------------------------------------
    with torch.no_grad():
        st, a = model._posterior_state(model.X)
        ell, s2 = model.kernel.length_scales.transform(), model.kernel.variance.transform()
        noise = model.likelihood.variance.transform()
        g = torch.Generator(device="cuda").manual_seed(7)          # same seed on every rank: v is replicated
        v = torch.randn(n, 1, dtype=torch.float64, device="cuda", generator=g)
        Xc = model.X
        u = torch.zeros_like(v)
        for j in st.mine:                                            # u += K(X, X_j) v_j  (+ noise v_j on the block's rows)
            c, wj = st.cols[j]
            Kj = nv.kern_fwd(0, Xc, Xc[c:c + wj], ell, s2)            # [n, wj]
            u += Kj @ v[c:c + wj]
            u[c:c + wj] += noise * v[c:c + wj]
            del Kj
        dist.all_reduce(u)
        z = torch.zeros_like(v)
        for j in st.mine:                                            # z = Ky^-1 u from the lower block triangle, mirrored
            c, wj = st.cols[j]
            s0 = st.slot[j] * st.w
            blk = st.A[c:, s0:s0 + wj]                                # rows >= c_j of block column j (diag block is full)
            z[c:] += blk @ u[c:c + wj]
            if blk.shape[0] > wj:
                z[c:c + wj] += blk[wj:].t() @ u[c + wj:]
        dist.all_reduce(z)
        out["inverse_identity_rel"] = float((z - v).abs().max() / v.abs().max())
        # a = Ky^-1 (y - m): the same inverse applied to the residual must reproduce the replicated a
        r = model.Y - model.mean_function(model.X)
        z2 = torch.zeros_like(r)
        for j in st.mine:
            c, wj = st.cols[j]
            s0 = st.slot[j] * st.w
            blk = st.A[c:, s0:s0 + wj]
            z2[c:] += blk @ r[c:c + wj]
            if blk.shape[0] > wj:
                z2[c:c + wj] += blk[wj:].t() @ r[c + wj:]
        dist.all_reduce(z2)
        out["alpha_rel"] = float((z2 - a).abs().max() / a.abs().max())
        del st
        model.__dict__.pop("_memo_store", None)
        torch.cuda.empty_cache()

    # ---- (2) analytic gradient vs central finite difference ------------------------------------------------------------
    loss = model.loss()
    loss.sum().backward()
    out["loss"] = float(loss.item())
    gd = torch.Generator().manual_seed(11)
    direction = [torch.randn(p.shape, generator=gd, dtype=torch.float64).to(p.device) for p in params]
    analytic = sum((p.grad * q).sum().item() for p, q in zip(params, direction))
    with torch.no_grad():
        for p, q in zip(params, direction):
            p.add_(args.eps * q)
        lp = model.loss().item()
        for p, q in zip(params, direction):
            p.sub_(2 * args.eps * q)
        lm = model.loss().item()
    fd = (lp - lm) / (2 * args.eps)
    out.update({"directional_derivative_analytic": analytic, "directional_derivative_fd": fd,
                "fd_rel": abs(fd - analytic) / abs(analytic)})
    if n == 131072:
        ref = -113952.01193680815
        out["loss_rel_vs_round1"] = abs(out["loss"] - ref) / abs(ref)
    ok = out["inverse_identity_rel"] < 1e-8 and out["alpha_rel"] < 1e-8 and out["fd_rel"] < 1e-5
    out["ok"] = bool(ok)
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()
    if not ok:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
