"""World-size-2 gloo tests (CPU) of the multi-process host logic: row sharding, the flat gradient all-reduce
used for data-parallel SVGP, `distribute()` bookkeeping and the sufficient-statistics identity the N-sharded
VFE relies on (sum of shard statistics == full statistics, checked with the oracle)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def _shard_and_reduce(rank, world):
    from gptorch_b200 import dist as gd, settings, kernels
    from gptorch_b200.models import VFE
    settings.set_default_device("cpu")
    n = 101
    lo, hi = gd.shard_rows(n, rank, world)
    rng = np.random.RandomState(0)
    X, Y = rng.rand(n, 2), rng.rand(n, 1)
    model = VFE(X[lo:hi], Y[lo:hi], kernels.Rbf(2), inducing_points=rng.rand(4, 2))
    model.distribute()
    # a fake local gradient: rank-dependent, summed by the flat all-reduce
    for i, p in enumerate(q for q in model.parameters() if q.requires_grad):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    gd.allreduce_grads(model)
    grads = [p.grad.flatten()[0].item() for p in model.parameters() if p.requires_grad]
    tmax = gd.max_over_ranks(10.0 + rank, torch.device("cpu"))
    return (lo, hi, model.num_data, grads, tmax)


def test_row_sharding_and_grad_allreduce():
    out = _run(_shard_and_reduce)
    (lo0, hi0, n0, g0, t0), (lo1, hi1, n1, g1, t1) = out
    assert (lo0, hi0, lo1, hi1) == (0, 51, 51, 101)
    assert n0 == n1 == 101                       # distribute(): global number of data
    assert g0 == g1 and g0[0] == 3.0 and g0[1] == 6.0   # (1 + 2) * (i + 1)
    assert t0 == t1 == 11.0


def _distribute_syncs_parameters(rank, world):
    """Every rank builds its model from its OWN shard and RNG state (different k-means centres, different q(u));
    distribute() must leave all replicas with rank 0's parameter values."""
    from gptorch_b200 import settings, kernels
    from gptorch_b200.models import VFE
    from gptorch_b200.param import Param
    settings.set_default_device("cpu")
    rng = np.random.RandomState(100 + rank)
    X, Y = rng.rand(60, 2), rng.rand(60, 1)
    np.random.seed(rank)
    model = VFE(X, Y, kernels.Matern32(2, length_scales=0.5 + rank, variance=1.0 + rank), num_inducing_points=5)
    model.extra = Param(torch.full((3,), float(rank + 7), dtype=torch.float64))
    before = model.Z.detach().clone()
    model.distribute()
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    return before.numpy(), flat.numpy()


def test_distribute_broadcasts_replicated_parameters():
    (z0, p0), (z1, p1) = _run(_distribute_syncs_parameters)
    assert not np.allclose(z0, z1)                  # the local initialisers really differ
    assert np.array_equal(p0, p1)                   # ... and distribute() removes the difference
    assert int(np.isclose(p0, 7.0).sum()) == 3 and not np.isclose(p0, 8.0).any()   # rank 0's values win


def test_shard_rows_partitions():
    from gptorch_b200.dist import shard_rows
    for n, w in ((10, 3), (7, 8), (1000003, 8), (0, 2)):
        spans = [shard_rows(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def _vfe_stats_identity(rank, world):
    """sum_r (A_r A_r^T, A_r Y_r) over row shards equals the full-data statistics: the VFE all-reduce point."""
    from oracle import gp_oracle as O
    from gptorch_b200.dist import shard_rows
    X, Y, g = O.synth_regression(400, 3)
    Z = O.synth_inducing(X, 12, g)
    h = O.Hyper("Matern32", [0.7, 0.9, 1.1], [1.3], [0.05])
    with torch.no_grad():
        L = O.chol(h.K(Z))
        lo, hi = shard_rows(400, rank, world)
        A = O.tri_solve(h.K(Z, X[lo:hi]), L)
        stats = torch.cat([(A @ A.t()).reshape(-1), (A @ Y[lo:hi]).reshape(-1), Y[lo:hi].pow(2).sum().reshape(1)])
        dist.all_reduce(stats)
        Afull = O.tri_solve(h.K(Z, X), L)
        full = torch.cat([(Afull @ Afull.t()).reshape(-1), (Afull @ Y).reshape(-1), Y.pow(2).sum().reshape(1)])
    return float((stats - full).abs().max() / full.abs().max())


def test_vfe_sufficient_statistics_shard_identity():
    errs = _run(_vfe_stats_identity)
    assert max(errs) < 1e-13


def test_block_column_cyclic_layout():
    """Host logic of the distributed Cholesky (gptorch_b200/models/dist_gpr.py): every block column has exactly one
    owner, slots are dense per rank, ragged last block."""
    from gptorch_b200.models.dist_gpr import block_columns, local_blocks, owner_of
    for n, panel, world in ((8300, 1024, 2), (131072, 2048, 8), (1000, 128, 3), (100, 128, 4)):
        cols = block_columns(n, panel)
        assert cols[0][0] == 0 and sum(w for _, w in cols) == n and all(w == panel for _, w in cols[:-1])
        seen = {}
        for r in range(world):
            mine, slot = local_blocks(n, panel, r, world)
            assert sorted(slot.values()) == list(range(len(mine)))
            for j in mine:
                assert owner_of(j, world) == r and j not in seen
                seen[j] = r
        assert sorted(seen) == list(range(len(cols)))
        loads = [sum(1 for j in seen if seen[j] == r) for r in range(world)]
        assert max(loads) - min(loads) <= 1


class _TorchOps:
    """Test-only stand-in for gptorch_b200.models.dist_gpr.NativeOps: the same primitive contract on torch CPU
    tensors (the oracle's arithmetic), so that the block-column-cyclic bookkeeping of the distributed Cholesky /
    inverse / gradient can run under gloo.  Lives in tests/ -- the product has no CPU path."""

    block = 1
    GEMM_NT, GEMM_TN, GEMM_NN = 0, 1, 2
    NAMES = {0: "Rbf", 1: "Exp", 2: "Matern32", 3: "Matern52"}

    def empty(self, rows, cols, device):
        return torch.full((rows, cols), float("nan"), dtype=torch.float64)

    def kern_fill(self, kind, Xr, Xc, ell, s2, out, ld):
        from oracle import gp_oracle as O
        out[:, : (Xr if Xc is None else Xc).shape[0]] = O.cov(self.NAMES[kind], Xr, Xc, ell, s2)

    def add_diag(self, blk, ld, value):
        blk.diagonal().add_(value.reshape(()))

    def factor_panel(self, blk, wp, ld):
        L = torch.linalg.cholesky(torch.tril(blk[:wp, :wp]) + torch.tril(blk[:wp, :wp], -1).t())
        blk[:wp, :wp] = L + torch.triu(torch.full_like(L, 7.0), 1)     # native potrf leaves scratch above the diagonal
        if blk.shape[0] > wp:
            blk[wp:, :wp] = torch.linalg.solve_triangular(L, blk[wp:, :wp].t().contiguous(), upper=False).t()
        return torch.zeros(1, dtype=torch.int32)

    def solve_lower(self, Lpp, xp):
        xp.copy_(torch.linalg.solve_triangular(torch.tril(Lpp), xp.contiguous(), upper=False))

    def logdet(self, Lpp):
        return Lpp.diagonal().log().sum()

    def sumsq(self, v):
        return v.pow(2).sum()

    def tri_inverse_t(self, Lpp):
        n = Lpp.shape[0]
        return torch.linalg.solve_triangular(torch.tril(Lpp), torch.eye(n, dtype=torch.float64), upper=False).t().contiguous()

    def gemm(self, mode, A, B, alpha=1.0, beta=0.0, C=None, flags=0):
        prod = A @ B.t() if mode == 0 else (A.t() @ B if mode == 1 else A @ B)
        if C is None:
            return alpha * prod
        C.copy_(alpha * prod + (beta * C if beta != 0.0 else 0.0))
        return C

    def gemv_t(self, A, Y, out):
        out.copy_(A.t() @ Y)
        return out

    def kern_bwd(self, kind, Xr, Xc, ell, s2, G):
        from oracle import gp_oracle as O
        e = ell.detach().clone().requires_grad_(True)
        s = s2.detach().clone().requires_grad_(True)
        with torch.enable_grad():
            (O.cov(self.NAMES[kind], Xr, Xc, e, s) * G).sum().backward()
        return e.grad, s.grad


def _dist_gpr_loss_and_grad(rank, world, n=203, panel=32, kind="Matern52"):
    from oracle import gp_oracle as O
    from gptorch_b200 import settings, kernels, likelihoods
    from gptorch_b200.models import DistributedGPR
    settings.set_default_device("cpu")
    d = 3
    X, Y, _ = O.synth_regression(n, d)
    ell, var, noise = [0.6, 0.9, 1.3], 1.7, 0.05
    kern = getattr(kernels, kind)(d, ARD=True, length_scales=np.array(ell), variance=var)
    model = DistributedGPR(X.numpy(), Y.numpy(), kern, likelihood=likelihoods.Gaussian(variance=noise), panel=panel,
                           ops=_TorchOps())
    loss = model.loss()
    loss.sum().backward()
    ref_loss, ref = O.gpr_loss_and_grads(kind, X, Y, ell, var, noise)
    got = {"variance": model.kernel.variance.grad, "length_scales": model.kernel.length_scales.grad,
           "noise": model.likelihood.variance.grad}
    rel = lambda a, b: float((a.reshape(-1) - b.reshape(-1)).abs().max() / b.abs().max())  # noqa: E731
    errs = {k: rel(got[k], ref[k]) for k in ref}
    # distributed prediction (mean, variance, full covariance) against the oracle's single-process posterior
    Xs = torch.rand(11, d, dtype=torch.float64, generator=torch.Generator().manual_seed(5))
    h = O.Hyper(kind, ell, var, noise)
    with torch.no_grad():
        mu_ref, var_ref = O.gpr_predict(h, X, Y, Xs, diag=True)
        _, cov_ref = O.gpr_predict(h, X, Y, Xs, diag=False)
    mu, v = model._predict(Xs, diag=True, block=4)            # several blocks of test points, ragged
    mu2, cov = model._predict(Xs, diag=False)
    errs["pred_mean"] = max(rel(mu, mu_ref), rel(mu2, mu_ref))
    errs["pred_var"] = float((v - var_ref).abs().max() / cov_ref.abs().max())
    errs["pred_cov"] = float((cov - cov_ref).abs().max() / cov_ref.abs().max())
    return (abs(loss.item() - ref_loss.item()) / abs(ref_loss.item()), errs)


@pytest.mark.parametrize("world,n,panel", [(2, 203, 32), (3, 130, 16), (2, 64, 64)])
def test_distributed_gpr_gradient_bookkeeping(world, n, panel):
    """Block-column-cyclic Cholesky -> T = L^-1 -> Ky^-1 = T^T T -> gradient reduction, on `world` gloo ranks with the
    torch stand-in for the native primitives, against the oracle's autograd (ragged last block, more ranks than
    blocks' worth of look-ahead, a single block)."""
    import functools
    out = _run(functools.partial(_dist_gpr_loss_and_grad, n=n, panel=panel), world=world)
    for loss_err, grad_err in out:
        assert loss_err < 1e-12
        assert max(grad_err.values()) < 1e-9, grad_err
