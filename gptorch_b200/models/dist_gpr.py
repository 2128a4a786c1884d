"""Exact GP regression whose covariance matrix does not fit one GPU: block-column-cyclic Cholesky, inverse and
gradient over the ranks of a torch.distributed group (BASELINE config #5: N = 131072 on 8 B200).

Layout.  The N x N matrix Ky = K(X, X) + noise * I is cut into block columns of `panel` columns; block column j lives
on rank j % R as a dense [N, panel] slab of that rank's local buffer (only rows >= j * panel are ever touched: the
lower triangle).  A rank's slabs are stored in ascending j, so "all my block columns <= p" is one contiguous column
range of the local buffer.  Every rank holds X (N x D, a few MB) and builds its own slabs with the fused covariance
kernel -- Ky itself is never communicated.

Loss (right-looking factorisation, one exchange per panel):
    owner of panel p : potrf of the panel's diagonal block, right-TRSM of the rows below      (local, native)
    all ranks        : ncclBroadcast of the factored panel  L[p*w:, p]   ((N - p*w) x w doubles)
    all ranks        : alpha-update  x_p = L_pp^-1 b_p ;  b_rest -= L[rest, p] x_p            (replicated, tiny)
    all ranks        : trailing update of their own block columns j > p with the received panel (NT GEMMs)
Look-ahead: the rank that owns panel p + 1 updates that block column first, factors it on a high-priority side stream
and starts its broadcast while everybody (itself included) is still applying panel p to the remaining columns.

Gradient (backward of the same node; the slabs are reused in place, L -> T = L^-1 -> Ky^-1):
    stage 1  T = L^-1, right-looking over block rows p (panel p of L is broadcast again):
                 T[p, j]  = -L_pp^-1 S[p, j]                     (owned j < p;  T[p, p] = L_pp^-1 on its owner)
                 S[i, j] +=  L[i, p] T[p, j]   for i > p         (one NN GEMM over all owned columns j <= p)
             S[i, j] lives where L[i, j] was: column j of L is dead once its panel has been sent.
             a = Ky^-1 r = T^T (L^-1 r) is then a local A^T y product per owned block column + one all-reduce.
    stage 2  Ky^-1 = T^T T, over block rows i (panel i of T is broadcast):
                 Kinv[i, j] = T[i:, i]^T T[i:, j]                (one TN GEMM over all owned columns j <= i)
             written over T[i, j], which later steps (rows > i only) no longer read.
    stage 3  W = 1/2 (dy Kinv - a a^T) is formed in place per owned slab and reduced against dK/d(ell, sigma2) by the
             dense covariance backward kernel (diagonal blocks weigh 1/2, blocks below the diagonal stand for both
             triangles); tr W gives the noise gradient; one all-reduce of D + 2 doubles.
Each stage is N^3/3 flop spread cyclically over the ranks, so a loss+grad evaluation is N^3 flop like the single-GPU
fused node (gptorch_b200/_autograd.py GPRLogLikFn), and every rank ends with the full hyper-parameter gradient.

The numerical primitives sit behind a small `ops` object (NativeOps: libgpb200.so) so that the block bookkeeping can
be exercised by the world-size-2 gloo tests with a torch stand-in that lives in tests/ -- the product has no CPU path.
"""
import math

import torch
import torch.distributed as dist
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import _autograd as ag
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


def owned_upto(p, rank, world, strict):
    """Number of block columns j owned by `rank` with j < p (strict) or j <= p."""
    last = p - 1 if strict else p
    return 0 if last < rank else (last - rank) // world + 1


class NativeOps:
    """The numerical primitives of the distributed evaluation on libgpb200.so (one GPU per rank)."""

    block = nv.NB
    GEMM_NT, GEMM_TN, GEMM_NN = nv.GEMM_NT, nv.GEMM_TN, nv.GEMM_NN

    def empty(self, rows, cols, device):
        return nv._aligned_empty(rows, cols, device)[0]

    def kern_fill(self, kind, Xr, Xc, ell, s2, out, ld):
        nv.kern_fwd(kind, Xr, Xc, ell, s2, out=out, ldk=ld)

    def add_diag(self, blk, ld, value):
        nv.add_diag_(blk, ld, value)

    def factor_panel(self, blk, wp, ld):
        """blk: rows c.. of one block column (h x >=wp view, leading dimension ld).  Diagonal block -> L_pp, rows
        below -> A L_pp^-T.  Returns the device info of the diagonal block's potrf."""
        h = blk.shape[0]
        dinv, info = nv.potrf_(blk[:wp], ld)
        if h > wp:
            nv.call("gpb_trsm_right_lt", nv.ptr(blk), wp, ld, nv.ptr(dinv), nv.ptr(blk[wp:]), h - wp, ld,
                    nv.stream_ptr())
        return info

    def solve_lower(self, Lpp, xp):
        nv.trsv_(Lpp, nv.tri_diag_inverse(Lpp), xp, False)

    def logdet(self, Lpp):
        return nv.logdet_sumsq(Lpp)[0]

    def sumsq(self, v):
        return nv.logdet_sumsq(None, v)[1]

    def tri_inverse_t(self, Lpp):
        """Dense upper-triangular L_pp^-T (zeros below the diagonal)."""
        Lc = nv._gemm_operand(Lpp)
        return ag._tinv(Lc, nv.tri_diag_inverse(Lc))

    SPLITK_MIN_K = 4096      # TN products with a long reduction and too few output tiles for 148 SMs are k-sliced

    def gemm(self, mode, A, B, alpha=1.0, beta=0.0, C=None, flags=0):
        if mode == nv.GEMM_TN and C is not None and alpha == 1.0 and beta == 0.0 and flags == 0:
            # Ky^-1 = T^T T, block row i: the output is only (panel width / 128) x (owned columns / 128) tiles while the
            # reduction runs over all rows below -- early block rows would occupy a fraction of the GPU for a long time
            k, m = A.shape
            n = B.shape[1]
            tiles = ((m + 127) // 128) * ((n + 127) // 128)
            if tiles < 2 * 148 and k >= self.SPLITK_MIN_K:
                splits = int(min(16, max(2, (4 * 148 + tiles - 1) // tiles)))
                kper = ((k + splits - 1) // splits + 15) // 16 * 16
                ldn = n + (n & 1)
                C3 = torch.zeros((splits, m, ldn), dtype=torch.float64, device=A.device)
                nv.gemm_splitk(mode, A, B, kper, C3, beta=0.0)
                torch.sum(C3[:, :, :n], dim=0, out=C) if C.is_contiguous() else C.copy_(C3[:, :, :n].sum(0))
                return C
        return nv.gemm(mode, A, B, alpha=alpha, beta=beta, C=C, flags=flags)

    def gemv_t(self, A, Y, out):
        """out = A^T Y (A tall, few right-hand sides)."""
        return nv.gemv_t(A, Y, out, beta=0.0)

    def kern_bwd(self, kind, Xr, Xc, ell, s2, G):
        g_ell, g_s2, _ = nv.kern_bwd(kind, Xr, Xc, ell, s2, G, False)
        return g_ell, g_s2


WAIT_LOG = None   # set to a list by bench.py to collect (start, end) events around every wait for a panel


class _Pipeline:
    """Double-buffered panel exchange: panel p + 1 is staged and broadcast on a side stream while the main stream
    works with panel p.  On a CPU device (gloo tests) everything is issued in order on the host."""

    def __init__(self, ops, n, w, device, group, side=None):
        self.bufs = [ops.empty(n, w, device)[:, :w] for _ in range(2)]
        self.cuda = device.type == "cuda"
        self.group = group
        self.ready = [None, None]
        self.done = [None, None]
        if self.cuda:
            self.main = torch.cuda.current_stream()
            self.side = side if side is not None else torch.cuda.Stream(priority=-1)

    def src(self, r):
        return dist.get_global_rank(self.group, r)

    def side_ctx(self, k, after_main):
        """Context in which buffer k may be refilled: waits until its last reader on the main stream has finished."""
        if not self.cuda:
            return _NullCtx()
        if self.done[k] is not None:
            self.side.wait_event(self.done[k])
        if after_main:
            self.side.wait_stream(self.main)
        return torch.cuda.stream(self.side)

    def mark_ready(self, k):
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record(self.side)
            self.ready[k] = ev

    def acquire(self, k):
        if self.cuda and self.ready[k] is not None:
            if WAIT_LOG is None:
                self.main.wait_event(self.ready[k])
            else:   # bench.py: how long the main stream stalls for the next panel (factor chain + broadcast)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(self.main)
                self.main.wait_event(self.ready[k])
                e1.record(self.main)
                WAIT_LOG.append((e0, e1))

    def release(self, k):
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record(self.main)
            self.done[k] = ev

    def finish(self):
        if self.cuda:
            self.main.wait_stream(self.side)

    def stream(self, cols, n, rank, world, stage, step):
        """Generic "for p: receive panel p, step(p, P)" loop with one panel of prefetch.  stage(p, dst) copies this
        rank's panel p (rows c_p.., it is the owner) into dst."""
        def fetch(p):
            c, wp = cols[p]
            k = p % 2
            with self.side_ctx(k, after_main=False):
                dst = self.bufs[k][: n - c]
                if owner_of(p, world) == rank:
                    stage(p, dst)
                dist.broadcast(dst, src=self.src(owner_of(p, world)), group=self.group)
                self.mark_ready(k)

        if self.cuda:
            self.side.wait_stream(self.main)
        fetch(0)
        for p in range(len(cols)):
            c, wp = cols[p]
            if p + 1 < len(cols):
                fetch(p + 1)
            self.acquire(p % 2)
            step(p, self.bufs[p % 2][: n - c])
            self.release(p % 2)
        self.finish()


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class _State:
    """What the backward pass needs from the forward pass (the slabs stay on the GPU that built them)."""
    __slots__ = ("A", "ld", "V", "cols", "mine", "slot", "n", "dy", "w", "rank", "world", "group", "kind", "ops", "side")


def _factorise(ops, kind, x, resid, ell, s2, noise, w, group, side=None):
    """Distributed Cholesky of Ky and the forward substitution V = L^-1 resid.  Returns (loglik [1], state)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n, dy = resid.shape
    dev = x.device
    cols = block_columns(n, w)
    mine, slot = local_blocks(n, w, rank, world)
    ld = max(len(mine), 1) * w

    # ---- this rank's slabs of Ky (lower part only) -----------------------------------------------------------
    A = ops.empty(n, ld, dev)[:, :ld]
    for j in mine:
        c, wj = cols[j]
        view = A[c:, slot[j] * w:]
        ops.kern_fill(kind, x[c:], x[c:c + wj], ell, s2, view, ld)
        ops.add_diag(view[:wj], ld, noise)

    b = ops.empty(n, dy, dev)[:, :dy]            # replicated right-hand side -> V = L^-1 resid (even row stride)
    b.copy_(resid)
    logdet = torch.zeros((), dtype=torch.float64, device=dev)
    info_total = torch.zeros(1, dtype=torch.int32, device=dev)
    pipe = _Pipeline(ops, n, w, dev, group, side)
    bufs = pipe.bufs

    def factor_panel(p):
        """Owner only: factor block column p in place and stage it (contiguous) in bufs[p % 2]."""
        c, wp = cols[p]
        blk = A[c:, slot[p] * w:]
        info_total.add_(ops.factor_panel(blk, wp, ld))
        bufs[p % 2][: n - c, :wp].copy_(blk[:, :wp])

    def update_column(j, p, P):
        """Block column j (local) -= contribution of the factored panel p held in P ((n - c_p) x w_p)."""
        c, wp = cols[p]
        cj, wj = cols[j]
        Cv = A[cj:, slot[j] * w: slot[j] * w + wj]
        ops.gemm(ops.GEMM_NT, P[cj - c:, :wp], P[cj - c: cj - c + wj, :wp], alpha=-1.0, beta=1.0, C=Cv)

    if owner_of(0, world) == rank:
        factor_panel(0)
    dist.broadcast(bufs[0], src=pipe.src(owner_of(0, world)), group=group)
    for p in range(len(cols)):
        c, wp = cols[p]
        P = bufs[p % 2][: n - c]
        pipe.acquire(p % 2)
        nxt = p + 1
        have_next = nxt < len(cols)
        i_own_next = have_next and owner_of(nxt, world) == rank
        # ---- look-ahead: bring panel p+1 up to date, factor it and broadcast it on the side stream -----------
        if have_next:
            k = nxt % 2
            with pipe.side_ctx(k, after_main=True):
                if i_own_next:
                    update_column(nxt, p, P)
                    factor_panel(nxt)
                dist.broadcast(bufs[k][: n - cols[nxt][0]], src=pipe.src(owner_of(nxt, world)), group=group)
                pipe.mark_ready(k)
        # ---- replicated alpha update and log-determinant with panel p ----------------------------------------
        Lpp = P[:wp, :wp]
        xp = b[c:c + wp]
        ops.solve_lower(Lpp, xp)
        logdet += ops.logdet(Lpp)
        if c + wp < n:
            ops.gemm(ops.GEMM_NN, P[wp:, :wp], xp, alpha=-1.0, beta=1.0, C=b[c + wp:])
        # ---- trailing update of this rank's remaining block columns ------------------------------------------
        for j in mine:
            if j > p and not (i_own_next and j == nxt):
                update_column(j, p, P)
        pipe.release(p % 2)
    pipe.finish()
    dist.all_reduce(info_total, op=dist.ReduceOp.MAX, group=group)
    if int(info_total.item()) != 0:
        raise torch.linalg.LinAlgError("distributed Cholesky: a diagonal block is not positive-definite")
    loglik = -0.5 * ops.sumsq(b) - dy * logdet - 0.5 * dy * n * math.log(2.0 * math.pi)

    st = _State()
    st.A, st.ld, st.V, st.cols, st.mine, st.slot = A, ld, b, cols, mine, slot
    st.n, st.dy, st.w, st.rank, st.world, st.group, st.kind, st.ops = n, dy, w, rank, world, group, kind, ops
    st.side = pipe.side if pipe.cuda else None
    return loglik.reshape(1), st


def _invert_factor(st):
    """Stage 1: the slabs (L, block-column cyclic) become T = L^-1 in place.  Returns a = T^T V = Ky^-1 resid."""
    ops, A, w, n, cols, slot, rank, world = st.ops, st.A, st.w, st.n, st.cols, st.slot, st.rank, st.world
    dev = A.device
    pipe = _Pipeline(ops, n, w, dev, st.group, st.side)
    tmp = ops.empty(w, st.ld, dev)[:, :st.ld]

    def stage(p, dst):
        c, wp = cols[p]
        dst[:, :wp].copy_(A[c:, slot[p] * w: slot[p] * w + wp])

    def step(p, P):
        c, wp = cols[p]
        U = ops.tri_inverse_t(P[:wp, :wp])                 # L_pp^-T, upper triangular (replicated, w x w)
        nb = owned_upto(p, rank, world, strict=True) * w   # columns of my block columns j < p
        mine_p = owner_of(p, world) == rank
        if nb:
            Srow = A[c:c + wp, :nb]
            t = tmp[:wp, :nb]
            t.copy_(Srow)
            ops.gemm(ops.GEMM_TN, U, t, alpha=-1.0, C=Srow, flags=nv.GF_KHI_M)      # T[p, j] = -L_pp^-1 S[p, j]
            if c + wp < n:
                ops.gemm(ops.GEMM_NN, P[wp:, :wp], Srow, beta=1.0, C=A[c + wp:, :nb])   # S[i, j] += L[i, p] T[p, j]
        if mine_p:
            s0 = slot[p] * w
            Tpp = A[c:c + wp, s0:s0 + wp]
            Tpp.copy_(U.t())                               # lower triangular with explicit zeros above the diagonal
            if c + wp < n:
                ops.gemm(ops.GEMM_NN, P[wp:, :wp], Tpp, C=A[c + wp:, s0:s0 + wp], flags=nv.GF_KLO_N)  # S[i,p] = L[i,p] T[p,p]

    pipe.stream(cols, n, rank, world, stage, step)

    a = torch.zeros_like(st.V)
    for j in st.mine:
        c, wj = cols[j]
        s0 = slot[j] * w
        ops.gemv_t(A[c:, s0:s0 + wj], st.V[c:], a[c:c + wj])
    dist.all_reduce(a, group=st.group)
    return a


def _inverse_from_t(st):
    """Stage 2: slabs T -> Ky^-1 = T^T T (lower block triangle; diagonal blocks full), in place."""
    ops, A, w, n, cols, slot, rank, world = st.ops, st.A, st.w, st.n, st.cols, st.slot, st.rank, st.world
    pipe = _Pipeline(ops, n, w, A.device, st.group, st.side)
    tmp = ops.empty(w, st.ld, A.device)[:, :st.ld]

    def stage(i, dst):
        c, wi = cols[i]
        dst[:, :wi].copy_(A[c:, slot[i] * w: slot[i] * w + wi])

    def step(i, Q):
        c, wi = cols[i]
        nb = owned_upto(i, rank, world, strict=False)
        if nb == 0:
            return
        last = rank + (nb - 1) * world                        # my last block column <= i
        ncol = (nb - 1) * w + cols[last][1]
        t = tmp[:wi, :ncol]
        ops.gemm(ops.GEMM_TN, Q[:, :wi], A[c:, :ncol], C=t)   # Kinv[i, j] = T[i:, i]^T T[i:, j],  j <= i
        A[c:c + wi, :ncol].copy_(t)

    pipe.stream(cols, n, rank, world, stage, step)


def _reduce_gradient(st, x, ell, s2, a):
    """Stage 3: (g_ell, g_sigma2, g_noise) = d(-loglik)/d(ell, sigma2, noise), summed over the ranks."""
    ops, A, w, cols, slot, dy = st.ops, st.A, st.w, st.cols, st.slot, st.dy
    g_ell = torch.zeros(ell.numel(), dtype=torch.float64, device=A.device)
    g_s2 = torch.zeros(1, dtype=torch.float64, device=A.device)
    g_noise = torch.zeros(1, dtype=torch.float64, device=A.device)
    for j in st.mine:
        c, wj = cols[j]
        s0 = slot[j] * w
        G = A[c:, s0:s0 + wj]
        ops.gemm(ops.GEMM_NT, a[c:], a[c:c + wj], alpha=-1.0, beta=float(dy), C=G)   # dy Kinv - a a^T  (= 2 W)
        ge, gs = ops.kern_bwd(st.kind, x[c:c + wj], x[c:c + wj], ell, s2, G[:wj])
        g_ell += 0.5 * ge
        g_s2 += 0.5 * gs
        g_noise += 0.5 * G[:wj].diagonal().sum()
        if G.shape[0] > wj:
            ge, gs = ops.kern_bwd(st.kind, x[c + wj:], x[c:c + wj], ell, s2, G[wj:])
            g_ell += ge
            g_s2 += gs
    flat = torch.cat([g_ell, g_s2, g_noise])
    dist.all_reduce(flat, group=st.group)
    return flat[: ell.numel()], flat[ell.numel(): ell.numel() + 1], flat[ell.numel() + 1:]


class DistGPRLogLikFn(Function):
    """GPR.log_likelihood (gptorch/models/gpr.py:47-67) over a process group; see the module docstring."""

    @staticmethod
    def forward(ctx, ops, kind, x, resid, ell, s2, noise, panel, group, side):
        with torch.no_grad(), nv.phase("dist_potrf"):
            loglik, st = _factorise(ops, kind, x, resid, ell, s2, noise, panel, group, side)
        ctx.st = st
        ctx.save_for_backward(x, ell, s2)
        return loglik

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        st, ctx.st = ctx.st, None
        if st is None:
            raise RuntimeError("DistGPRLogLikFn: the factor slabs were consumed by a previous backward pass")
        x, ell, s2 = ctx.saved_tensors
        with nv.phase("dist_trtri"):
            a = _invert_factor(st)
        with nv.phase("dist_lauum"):
            _inverse_from_t(st)
        with nv.phase("dist_grad"):
            g_ell, g_s2, g_noise = _reduce_gradient(st, x, ell.reshape(-1), s2, a)
        g = g.reshape(())
        need = ctx.needs_input_grad
        return (None, None, None,
                (-g) * a if need[3] else None,
                (-g) * g_ell.reshape(ell.shape) if need[4] else None,
                (-g) * g_s2.reshape(s2.shape) if need[5] else None,
                (-g) * g_noise.reshape(-1) if need[6] else None,
                None, None, None)


class DistributedGPR(GPR):
    """GPR whose log_likelihood() (and its gradient) runs block-column cyclic over `group`.

    Every rank constructs the model with the SAME (x, y) and hyper-parameters and calls loss() / backward()
    collectively; every rank ends up with the same loss and the same, complete gradients.  `panel` is the block-column
    width: the owner's per-panel chain (column update, potrf, TRSM, broadcast) is serial and grows with the width,
    the trailing GEMMs want a long k -- 1024 is the measured optimum on 8 B200 at N = 131072 (3.28 s per
    log-likelihood; 512: 3.34 s, 2048: 3.55 s).
    """

    def __init__(self, x, y, kernel, mean_function=None, likelihood=None, group=None, panel=1024, name="dist_gpr",
                 ops=None):
        super().__init__(x, y, kernel, mean_function=mean_function, likelihood=likelihood, name=name)
        self._ops = ops if ops is not None else NativeOps()
        if panel % self._ops.block != 0:
            raise ValueError("panel must be a multiple of %d" % self._ops.block)
        self._group = group if group is not None else dist.group.WORLD
        self._panel = panel
        self._side = None

    def _kind(self):
        kind = _native_kind(self.kernel)
        if kind is None or not isinstance(self.likelihood, Gaussian):
            raise NotImplementedError("DistributedGPR needs a stationary kernel and the Gaussian likelihood")
        return kind

    def _posterior_state(self, x):
        """(slabs holding Ky^-1 as a lower block triangle, a = Ky^-1 (Y - m(x))) -- what every prediction needs.
        Built by the factorisation and stages 1-2 of the gradient, cached per parameter / data state (GPModel._memo)."""
        def compute():
            xc = nv._c(x.to(torch.float64))
            _, st = _factorise(self._ops, self._kind(), xc, self.Y - self.mean_function(xc),
                               self.kernel.length_scales.transform(), self.kernel.variance.transform(),
                               self.likelihood.variance.transform(), self._panel, self._group, self._side)
            a = _invert_factor(st)
            _inverse_from_t(st)
            return st, a

        return self._memo("dist_posterior", x, compute)

    @torch.no_grad()
    def _predict(self, x_new, diag=True, x=None, block=1024):
        """p(f* | y) (gptorch/models/gpr.py:88-117) with the training covariance spread over the ranks: the mean is
        K(x*, X) a with the replicated a = Ky^-1 r; the quadratic form k*^T Ky^-1 k* is summed slab by slab from the
        distributed inverse (diagonal blocks once, blocks below the diagonal twice) and all-reduced.  Test points are
        processed in blocks so that K(X, x*) never exceeds N x `block`.  Collective; not differentiable."""
        x = x if x is not None else self.X
        if self._side is None and x.is_cuda:
            self._side = torch.cuda.Stream(priority=-1)
        st, a = self._posterior_state(x)
        ops, kind = self._ops, st.kind
        ell, s2 = self.kernel.length_scales.transform(), self.kernel.variance.transform()
        xs_all = nv._c(x_new.to(torch.float64))
        xc = nv._c(x.to(torch.float64))
        ns, dy, w = xs_all.shape[0], st.dy, st.w
        dev = xc.device
        mean = torch.zeros((ns, dy), dtype=torch.float64, device=dev)
        quad = torch.zeros(ns if diag else (ns, ns), dtype=torch.float64, device=dev)
        if not diag:
            block = max(ns, 1)              # the full covariance needs every pair of test points at once
        for b0 in range(0, ns, block):
            xs = xs_all[b0: b0 + block]
            nb = xs.shape[0]
            for j in st.mine:
                c, wj = st.cols[j]
                s0 = st.slot[j] * w
                Ks = ops.empty(st.n - c, nb, dev)[:, :nb]
                ops.kern_fill(kind, xc[c:], xs, ell, s2, Ks, Ks.stride(0))          # K(X[c:], x*)
                Kj = Ks[:wj].clone()
                # this rank's share of the mean: block column j of K(x*, X) times a_j
                mean[b0: b0 + nb] += ops.gemm(ops.GEMM_TN, Kj, a[c:c + wj])
                Ks[:wj].mul_(0.5)
                Z = ops.gemm(ops.GEMM_TN, st.A[c:, s0:s0 + wj], Ks)                  # (wj x nb)
                if diag:
                    quad[b0: b0 + nb] += 2.0 * (Z * Kj).sum(0)
                else:
                    S = ops.gemm(ops.GEMM_TN, Z, Kj)
                    quad += S + S.t()
        flat = torch.cat([mean.reshape(-1), quad.reshape(-1)])
        dist.all_reduce(flat, group=self._group)
        mean = flat[: ns * dy].reshape(ns, dy) + self.mean_function(xs_all)
        quad = flat[ns * dy:].reshape(quad.shape)
        if diag:
            var = (s2.reshape(()) - quad)[:, None].expand(ns, dy)
        else:
            Kss = ops.empty(ns, ns, dev)[:, :ns]
            ops.kern_fill(kind, xs_all, None, ell, s2, Kss, Kss.stride(0))
            var = Kss - quad
        return mean, var

    def log_likelihood(self, x=None, y=None):
        x = x if x is not None else self.X
        y = y if y is not None else self.Y
        if x.shape[0] != y.shape[0]:
            raise ValueError("X and Y must have same # data.")
        kind = self._kind()
        if x.is_cuda and self._side is None:
            self._side = torch.cuda.Stream(priority=-1)
        x = nv._c(x.to(torch.float64))
        resid = y - self.mean_function(x)
        return DistGPRLogLikFn.apply(self._ops, kind, x, resid, self.kernel.length_scales.transform(),
                                     self.kernel.variance.transform(), self.likelihood.variance.transform(),
                                     self._panel, self._group, self._side)
