"""Exact GP regression whose covariance matrix does not fit one GPU: block-column-cyclic Cholesky over the ranks of
a torch.distributed group (BASELINE config #5: N = 131072 on 8 B200).

Layout.  The N x N matrix Ky = K(X, X) + noise * I is cut into block columns of `panel` columns; block column j lives
on rank j % R as a dense [N, panel] slab of that rank's local buffer (only rows >= j * panel are ever touched: the
lower triangle).  Every rank holds X (N x D, a few MB) and builds its own slabs with the fused covariance kernel --
Ky itself is never communicated.

Factorisation (right-looking, one exchange per panel):
    owner of panel p : potrf of the panel's diagonal block, right-TRSM of the rows below      (local, native)
    all ranks        : ncclBroadcast of the factored panel  L[p*w:, p]   ((N - p*w) x w doubles)
    all ranks        : alpha-update  x_p = L_pp^-1 b_p ;  b_rest -= L[rest, p] x_p            (replicated, tiny)
    all ranks        : trailing update of their own block columns j > p with the received panel (NT GEMMs)
Look-ahead: the rank that owns panel p + 1 updates that block column first, factors it on a high-priority side stream
and starts its broadcast while everybody (itself included) is still applying panel p to the remaining columns.

What is computed: log p(y | X, theta) (GPR.log_likelihood, gptorch/models/gpr.py:47-67) -- loss only.  The gradient
needs a distributed (L L^T)^-1 and is not implemented in this round ("acceptable first milestone").
"""
import math

import torch
import torch.distributed as dist

from .. import _native as nv
from ..likelihoods import Gaussian
from .gpr import GPR, _native_kind


def block_columns(n, panel):
    """[(start, width)] of the block columns of an n x n matrix."""
    return [(c, min(panel, n - c)) for c in range(0, n, panel)]


def owner_of(j, world):
    return j % world


def local_blocks(n, panel, rank, world):
    """Global block-column indices stored on `rank` (cyclic) and their slot in the local buffer."""
    cols = block_columns(n, panel)
    mine = [j for j in range(len(cols)) if owner_of(j, world) == rank]
    return mine, {j: s for s, j in enumerate(mine)}


class DistributedGPR(GPR):
    """GPR whose log_likelihood() runs the block-column-cyclic factorisation over `group`.

    Every rank constructs the model with the SAME (x, y) and hyper-parameters and calls loss() collectively.
    """

    def __init__(self, x, y, kernel, mean_function=None, likelihood=None, group=None, panel=2048, name="dist_gpr"):
        super().__init__(x, y, kernel, mean_function=mean_function, likelihood=likelihood, name=name)
        if panel % nv.NB != 0:
            raise ValueError("panel must be a multiple of %d" % nv.NB)
        self._group = group if group is not None else dist.group.WORLD
        self._panel = panel
        self._side = None

    @torch.no_grad()
    def log_likelihood(self, x=None, y=None):
        x = x if x is not None else self.X
        y = y if y is not None else self.Y
        if x.shape[0] != y.shape[0]:
            raise ValueError("X and Y must have same # data.")
        kind = _native_kind(self.kernel)
        if kind is None or not isinstance(self.likelihood, Gaussian):
            raise NotImplementedError("DistributedGPR needs a stationary kernel and the Gaussian likelihood")
        group, w = self._group, self._panel
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        src_rank = lambda r: dist.get_global_rank(group, r)  # noqa: E731
        n, dy = y.shape
        dev = x.device
        x = nv._c(x.to(torch.float64))
        ell = self.kernel.length_scales.transform()
        s2 = self.kernel.variance.transform()
        noise = self.likelihood.variance.transform()
        cols = block_columns(n, w)
        mine, slot = local_blocks(n, w, rank, world)

        # ---- build this rank's slabs of Ky (lower part only) ----------------------------------------------
        ld = max(len(mine), 1) * w
        A = torch.empty((n, ld), dtype=torch.float64, device=dev)
        for j in mine:
            c, wj = cols[j]
            view = A[c:, slot[j] * w:]
            nv.kern_fwd(kind, x[c:], x[c:c + wj], ell, s2, out=view, ldk=ld)
            nv.add_diag_(view[:wj], ld, noise)

        bbuf, _ = nv._aligned_empty(n, dy, dev)                    # replicated right-hand side -> alpha (even row stride)
        b = bbuf[:, :dy]
        b.copy_(y - self.mean_function(x))
        logdet = torch.zeros((), dtype=torch.float64, device=dev)
        bufs = [torch.empty((n, w), dtype=torch.float64, device=dev) for _ in range(2)]
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream(priority=-1)
        side = self._side
        info_total = torch.zeros(1, dtype=torch.int32, device=dev)

        def factor_panel(p):
            """Owner only: factor block column p in place and stage it (contiguous) in bufs[p % 2]."""
            c, wp = cols[p]
            blk = A[c:, slot[p] * w:]
            dinv, info = nv.potrf_(blk[:wp], ld)
            info_total.add_(info)
            if c + wp < n:
                nv.call("gpb_trsm_right_lt", nv.ptr(blk), wp, ld, nv.ptr(dinv), nv.ptr(blk[wp:]), n - c - wp, ld,
                        nv.stream_ptr())
            bufs[p % 2][: n - c, :wp].copy_(blk[:, :wp])

        def update_column(j, p, P):
            """Block column j (local) -= contribution of the factored panel p held in P ((n - c_p) x w_p)."""
            c, wp = cols[p]
            cj, wj = cols[j]
            Cv = A[cj:, slot[j] * w: slot[j] * w + wj]
            nv.gemm(nv.GEMM_NT, P[cj - c:, :wp], P[cj - c: cj - c + wj, :wp], alpha=-1.0, beta=1.0, C=Cv)

        done_with = [None, None]      # event: main-stream work reading bufs[k] has been issued and finished
        ready = [None, None]          # event: bufs[k] holds the broadcast panel
        # panel 0
        if owner_of(0, world) == rank:
            factor_panel(0)
        dist.broadcast(bufs[0], src=src_rank(owner_of(0, world)), group=group)
        for p in range(len(cols)):
            c, wp = cols[p]
            P = bufs[p % 2][: n - c]
            if ready[p % 2] is not None:
                main.wait_event(ready[p % 2])
            nxt = p + 1
            have_next = nxt < len(cols)
            i_own_next = have_next and owner_of(nxt, world) == rank
            # ---- look-ahead: bring panel p+1 up to date, factor it and broadcast it on the side stream -------
            if have_next:
                k = nxt % 2
                if done_with[k] is not None:
                    side.wait_event(done_with[k])          # bufs[k] (panel p-1) is no longer being read
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    if i_own_next:
                        update_column(nxt, p, P)
                        factor_panel(nxt)
                    dist.broadcast(bufs[k], src=src_rank(owner_of(nxt, world)), group=group)
                    ev = torch.cuda.Event()
                    ev.record(side)
                    ready[k] = ev
            # ---- replicated alpha update and log-determinant with panel p -------------------------------------
            Lpp = P[:wp, :wp]
            dinv_p = nv.tri_diag_inverse(Lpp)
            xp = b[c:c + wp]
            nv.trsv_(Lpp, dinv_p, xp, False)
            logdet += nv.logdet_sumsq(Lpp)[0]
            if c + wp < n:
                nv.gemm(nv.GEMM_NN, P[wp:, :wp], xp, alpha=-1.0, beta=1.0, C=b[c + wp:])
            # ---- trailing update of this rank's remaining block columns ---------------------------------------
            for j in mine:
                if j > p and not (i_own_next and j == nxt):
                    update_column(j, p, P)
            ev = torch.cuda.Event()
            ev.record(main)
            done_with[p % 2] = ev
        main.wait_stream(side)
        dist.all_reduce(info_total, op=dist.ReduceOp.MAX, group=group)
        if int(info_total.item()) != 0:
            raise torch.linalg.LinAlgError("distributed Cholesky: a diagonal block is not positive-definite")
        sumsq = nv.logdet_sumsq(None, b)[1]
        loglik = -0.5 * sumsq - dy * logdet - 0.5 * dy * n * math.log(2.0 * math.pi)
        return loglik.reshape(1)
