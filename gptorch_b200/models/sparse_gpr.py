"""Sparse GPs with inducing points: VFE (Titsias 2009) and SVGP (Hensman et al. 2013/2015)
(reference: gptorch/models/sparse_gpr.py).

Both models work on ROW panels Kfu = K(x, Z) [n, M] (the transpose of the reference's Kuf) so that every
O(n M^2) step is an NT / TN product on the native DMMA engine: A^T = Kfu L^-T, A A^T = (A^T)^T A^T, ...
VFE never materialises Kuf: it streams row chunks of x through `VfeStatsFn`, which is also the point where an
N-sharded multi-GPU run all-reduces its M x M statistics.
"""
import numpy as np
import torch
from torch.distributions.transforms import LowerCholeskyTransform

from .. import _autograd as ag
from .. import settings
from ..functions import cholesky, trtrs, mm, mm_nt, mm_tn
from ..likelihoods import Gaussian
from ..mean_functions import Zero
from ..model import Param
from ..util import KMEANS_HOST_MAX_ROWS, as_tensor, kmeans_centers, kmeans_centers_device, torch_dtype
from .base import GPModel
from .gpr import GPR, _native_kind

MINIBATCH_PERMUTE_MAX = 1 << 22   # above this many rows the minibatch indices are drawn in O(batch)
VFE_CHUNK_ROWS = 1 << 17   # rows of x per streamed panel (128Ki x M fp64: 1 GiB at M = 1024)


class _InducingPointsGP(GPModel):
    """Common part of the inducing-point models: Z is a trainable, un-transformed Param
    (gptorch/models/sparse_gpr.py:24-73)."""

    def __init__(self, x, y, kernel, num_inducing_points=None, inducing_points=None, mean_function=None,
                 likelihood=None, data_on_host=False):
        super().__init__(x, y, kernel, likelihood, mean_function, data_on_host=data_on_host)
        if inducing_points is None:
            if num_inducing_points is None:
                num_inducing_points = np.clip(x.shape[0] // 10, 1, 100)
            if x.shape[0] > KMEANS_HOST_MAX_ROWS and torch.cuda.is_available():
                inducing_points = kmeans_centers_device(x, num_inducing_points)
            else:
                x_host = x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else x
                inducing_points = kmeans_centers(x_host, num_inducing_points, perturb_if_fail=True)
        self.Z = Param(as_tensor(inducing_points))
        self._group = None          # torch.distributed process group when the rows of X are sharded
        self._num_data_global = None

    @property
    def num_inducing(self):
        return self.Z.shape[0]

    def distribute(self, group=None):
        """Declare that self.X / self.Y hold this rank's shard of the rows.  Must be called on
        every rank after torch.distributed.init_process_group and before the first loss().

        Every parameter (kernel, likelihood, Z, q(u)) is REPLICATED: the summed statistics / gradients are only the
        global bound if all ranks evaluate them at the same values.  The initialisers depend on the local shard and
        the local RNG (k-means for Z, SVGP._init_posterior), so the values of the group's first rank are broadcast
        here; afterwards identical optimiser steps on the all-reduced gradients keep the replicas equal."""
        import torch.distributed as dist
        self._group = group if group is not None else dist.group.WORLD
        src = dist.get_global_rank(self._group, 0)
        for p in self.parameters():
            dist.broadcast(p.data, src=src, group=self._group)
        self.__dict__.pop("_memo_store", None)
        n = torch.tensor([float(self.Y.shape[0])], dtype=torch_dtype, device=self.compute_device)
        dist.all_reduce(n, group=self._group)
        self._num_data_global = int(n.item())
        return self

    @property
    def num_data(self):
        return self._num_data_global if self._num_data_global is not None else self.Y.shape[0]


class FITC(_InducingPointsGP):
    """Placeholder, as in the reference (gptorch/models/sparse_gpr.py:76-90)."""
    pass


class VFE(_InducingPointsGP):
    """Variational free energy (collapsed bound) sparse GP regression."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        assert isinstance(self.mean_function, Zero), "Mean functions not implemented for VFE yet."

    def _stats(self, x, L):
        """(A A^T, A Y, sum Kdiag, sum Y^2) with A = L^-1 Kuf and Y = self.Y (the reference ignores the y
        argument of log_likelihood, gptorch/models/sparse_gpr.py:125)."""
        kind = _native_kind(self.kernel)
        if kind is not None:
            return ag.VfeStatsFn.apply(kind, x, self.Y, self.Z, self.kernel.length_scales.transform(),
                                       self.kernel.variance.transform(), L, VFE_CHUNK_ROWS, self._group,
                                       self._use_phi_form(L, x.shape[0]))
        if self._group is not None:
            raise NotImplementedError("row-sharded VFE needs a stationary kernel")
        At = ag.TrsmRightFn.apply(self.kernel.K(x, self.Z), L, ag._dinv_of(L))
        return mm_tn(At, At), mm_tn(At, self.Y), self.kernel.Kdiag(x).sum(), self.Y.pow(2).sum()

    def _use_phi_form(self, L, rows):
        """settings.vfe_phi_form: the 3 N M^2 Phi form when Kuu is well enough conditioned (one host read per
        evaluation; every rank sees the same replicated L and therefore takes the same branch).  "auto" keeps the
        reference order for short data sets (nothing to gain below 16 M rows per rank; with row sharding all ranks must
        agree, so the local row count only switches the form off when every rank is short) and while the evaluation is
        being captured into a CUDA graph (no host reads there)."""
        mode = settings.vfe_phi_form
        if mode is True or mode is False:
            return mode
        if torch.cuda.is_current_stream_capturing() or (self._group is None and rows < 16 * L.shape[0]):
            return False
        with torch.no_grad():
            Lc = ag.nv._gemm_operand(L.detach())
            cond = ag.kuu_condition_estimate(Lc, ag._tinv(Lc, ag._dinv_of(L)))
        self.last_kuu_condition = cond
        return cond <= settings.vfe_phi_cond_max

    def _core(self, x):
        noise = self.likelihood.variance.transform()
        m = self.num_inducing
        L = cholesky(self.kernel.K(self.Z))
        AA, AY, kd_sum, yy = self._stats(x, L)
        AAT = AA / noise
        B = AAT + torch.eye(m, dtype=torch_dtype, device=AAT.device)
        LB = cholesky(B)
        c = trtrs(AY, LB) / noise
        return noise, L, AAT, LB, c, kd_sum, yy

    def log_likelihood(self, x=None, y=None):
        """Titsias' bound, eq. (9); 0-dim tensor (gptorch/models/sparse_gpr.py:108-153)."""
        x = x if x is not None else self.X
        y = y if y is not None else self.Y
        if x.shape[0] != y.shape[0]:
            raise ValueError("X and Y must have same # data.")
        n = self.num_data if self._group is not None else x.shape[0]
        dy = self.output_dimension
        noise, L, AAT, LB, c, kd_sum, yy = self._core(x)
        elbo = -0.5 * dy * n * np.log(2 * np.pi)
        elbo = elbo - dy * ag.LogDetFn.apply(LB)
        elbo = elbo - 0.5 * dy * n * noise.log()
        elbo = elbo - 0.5 * (yy + dy * kd_sum) / noise
        elbo = elbo + 0.5 * c.pow(2).sum()
        elbo = elbo + 0.5 * dy * AAT.diagonal().sum()
        return elbo[0]

    def _predict(self, x_new, diag=True, x=None):
        """p(f* | y) with the inducing outputs integrated out (gptorch/models/sparse_gpr.py:155-195).
        Unlike the reference this does not freeze Z as a side effect."""
        x = x if x is not None else self.X
        # O(N M^2) part: streamed once per parameter state under no_grad (the reference recomputes it per call)
        noise, L, AAT, LB, c, _, _ = self._memo("core", x, lambda: self._core(x))
        T1 = ag.TrsmRightFn.apply(self.kernel.K(x_new, self.Z), L, ag._dinv_of(L))     # (L^-1 Kus)^T
        T2 = ag.TrsmRightFn.apply(T1, LB, ag._dinv_of(LB))                             # (LB^-1 L^-1 Kus)^T
        mean = mm(T2, c)
        if diag:
            var = (self.kernel.Kdiag(x_new) - ag.RowSumSqFn.apply(T1) + ag.RowSumSqFn.apply(T2))[:, None].expand_as(mean)
        else:
            var = self.kernel.K(x_new) + mm_nt(T2, T2) - mm_nt(T1, T1)
        return mean, var


def draw_minibatch_indices(n, batch_size):
    """Indices of a minibatch drawn without replacement from n rows (numpy int64)."""
    if n <= MINIBATCH_PERMUTE_MAX:
        return np.random.permutation(n)[:batch_size]        # the reference's draw, same RNG stream
    # O(batch) draw without replacement: a full permutation of 1e8 indices per step is 0.8 GB of host work
    return np.random.default_rng(np.random.randint(1 << 31)).choice(n, size=batch_size, replace=False)


class HostBatchStream:
    """Minibatches of a data set that stays in page-locked HOST memory (N = 1e8 rows of D = 32 are
    25.6 GB; a data set beyond HBM cannot be indexed on the device like gptorch/models/sparse_gpr.py:209-211 does).

    next() hands out the batch whose rows were gathered into a pinned staging buffer and copied to the device on a
    side stream WHILE the previous step was computing, then starts gathering the following batch on a worker thread
    (torch.index_select releases the GIL).  Two staging buffers per tensor; a CUDA event per buffer keeps the worker
    from overwriting a buffer whose host-to-device copy is still in flight.  Indices come from the same host RNG
    stream as the resident path, one draw per batch, in order.
    """

    def __init__(self, X, Y, batch_size, device):
        import threading
        self.X, self.Y, self.b, self.dev = X, Y, int(batch_size), device
        pin = torch.cuda.is_available()
        self.stage = [(torch.empty((self.b, X.shape[1]), dtype=X.dtype, pin_memory=pin),
                       torch.empty((self.b, Y.shape[1]), dtype=Y.dtype, pin_memory=pin)) for _ in range(2)]
        self.copied = [None, None]            # event: the H2D copy out of staging buffer k has completed
        self.stream = torch.cuda.Stream(device=device) if device.type == "cuda" else None
        self._threading = threading
        self.k = 0
        self.pending = None
        self._start()

    def _gather(self, k, idx):
        if self.copied[k] is not None:
            self.copied[k].synchronize()
        sx, sy = self.stage[k]
        torch.index_select(self.X, 0, idx, out=sx)
        torch.index_select(self.Y, 0, idx, out=sy)

    def _start(self):
        idx = torch.from_numpy(np.ascontiguousarray(draw_minibatch_indices(self.Y.shape[0], self.b), dtype=np.int64))
        th = self._threading.Thread(target=self._gather, args=(self.k, idx), daemon=True)
        th.start()
        self.pending = (self.k, th)
        self.k ^= 1

    def next(self):
        k, th = self.pending
        th.join()
        sx, sy = self.stage[k]
        if self.stream is None:
            x, y = sx.clone(), sy.clone()
        else:
            cur = torch.cuda.current_stream(self.dev)
            with torch.cuda.stream(self.stream):
                x = sx.to(self.dev, non_blocking=True)
                y = sy.to(self.dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.stream)
            self.copied[k] = ev
            cur.wait_stream(self.stream)
            x.record_stream(cur)
            y.record_stream(cur)
        self._start()
        return x, y


def minibatch(loss_func):
    """Draw a random subset of the data when none is given (gptorch/models/sparse_gpr.py:198-216)."""

    def wrapped(obj, x=None, y=None):
        if x is not None:
            assert y is not None
        elif obj.batch_size is not None:
            if obj.data_on_host:
                stream = obj.__dict__.get("_host_stream")
                if stream is None or stream.b != obj.batch_size or stream.X is not obj.X or stream.Y is not obj.Y:
                    stream = obj.__dict__["_host_stream"] = HostBatchStream(obj.X, obj.Y, obj.batch_size, obj.compute_device)
                x, y = stream.next()
            else:
                i = torch.as_tensor(draw_minibatch_indices(obj.Y.shape[0], obj.batch_size), device=obj.X.device)
                x, y = obj.X[i, :], obj.Y[i, :]
        elif obj.data_on_host:
            raise ValueError("a model whose data stay on the host needs batch_size (or explicit x, y)")
        else:
            x, y = obj.X, obj.Y
        return loss_func(obj, x, y)

    return wrapped


class SVGP(_InducingPointsGP):
    """Sparse variational GP with an explicit Gaussian q(u) = N(induced_output_mean + m(Z), L_S L_S^T)."""

    def __init__(self, y, x, kernel, num_inducing_points=None, inducing_points=None, mean_function=None,
                 likelihood=None, batch_size=None, data_on_host=False):
        # NB the first two positional arguments are (inputs, outputs) despite their names -- the reference
        # swaps the names but passes them through positionally (gptorch/models/sparse_gpr.py:230-253).
        if likelihood is None:
            likelihood = Gaussian()
        super().__init__(y, x, kernel, num_inducing_points=num_inducing_points, inducing_points=inducing_points,
                         mean_function=mean_function, likelihood=likelihood, data_on_host=data_on_host)
        self.batch_size = batch_size
        self.induced_output_mean, self.induced_output_chol_cov = self._init_posterior()

    def _whitened(self, chol_kuu):
        """beta = L^-1 L_S and t = L^-1 m_u, shared by the bound and the prediction."""
        L_S = self.induced_output_chol_cov.transform()
        return trtrs(L_S, chol_kuu), trtrs(self.induced_output_mean, chol_kuu), L_S

    @minibatch
    def log_likelihood(self, x, y):
        """Variational bound on a (mini)batch; 0-dim tensor (gptorch/models/sparse_gpr.py:263-308)."""
        if x.shape[0] != y.shape[0]:
            raise ValueError("X and Y must have same # data.")
        chol_kuu = cholesky(self.kernel.K(self.Z))
        beta, t, L_S = self._whitened(chol_kuu)
        if self._use_quadratic_form(chol_kuu, x.shape[0]):
            f_mean, f_var = self._moments_quadratic(x, chol_kuu, beta, t)
        else:
            f_mean, f_var = self._predict(x, diag=True, chol_kuu=chol_kuu, _whitened=(beta, t))
        if isinstance(self.likelihood, Gaussian):
            mll = torch.stack([self.likelihood.expected_log_density(m_i, v_i, y_i)
                               for m_i, v_i, y_i in zip(f_mean.t(), f_var.t(), y.t())]).sum()
        else:   # the abstract Likelihood API, as the reference calls it (gptorch/models/sparse_gpr.py:274-281)
            mll = torch.stack([self.likelihood.propagate_log(torch.distributions.Normal(m_i, torch.sqrt(v_i)), y_i)
                               for m_i, v_i, y_i in zip(f_mean.t(), f_var.t(), y.t())]).sum()
        batch = x.shape[0]
        world = 1
        if self._group is not None:
            import torch.distributed as dist
            world = dist.get_world_size(self._group)
            batch = batch * world           # every rank draws the same batch size from its shard
        mll = mll * (self.num_data / batch)
        # KL(q(u) || p(u)), summed over output dimensions (the mean function shifts both and cancels):
        m, dy = self.num_inducing, self.output_dimension
        kl = 0.5 * dy * (beta.pow(2).sum() - m + 2.0 * ag.LogDetFn.apply(chol_kuu) - 2.0 * L_S.diagonal().log().sum())
        kl = kl + 0.5 * t.pow(2).sum()
        return mll - kl / world             # summed over ranks this is the global bound

    def _use_quadratic_form(self, chol_kuu, rows):
        """settings.svgp_quadratic_form: do the M x M algebra first (SvgpMomentsFn, 3 B M^2 flop per loss+grad instead
        of 6 B M^2) when the batch is tall and Kuu is well enough conditioned -- the same trade and the same gate as the
        VFE Phi form (settings.vfe_phi_cond_max)."""
        mode = settings.svgp_quadratic_form
        if mode is True or mode is False:
            return mode
        if (torch.cuda.is_current_stream_capturing() or rows < 16 * self.num_inducing
                or not isinstance(self.likelihood, Gaussian)):
            return False
        with torch.no_grad():
            Lc = ag.nv._gemm_operand(chol_kuu.detach())
            cond = ag.kuu_condition_estimate(Lc, ag._tinv(Lc, ag._dinv_of(chol_kuu)))
        self.last_kuu_condition = cond
        return cond <= settings.vfe_phi_cond_max

    def _moments_quadratic(self, x, chol_kuu, beta, t):
        """q(f) mean and variance on a batch with the M x M algebra first:
        C = Kuu^-1 (S - Kuu) Kuu^-1 = (T beta)(T beta)^T - T T^T and m = Kuu^-1 m_u = T t with T = L^-T."""
        T = ag.TriInvTFn.apply(chol_kuu, ag._dinv_of(chol_kuu))
        U = mm(T, beta, b_lower=True)
        C = ag.SyrkFn.apply(U) - ag.SyrkFn.apply(T, True)
        mvec = mm(T, t)
        mean, q = ag.SvgpMomentsFn.apply(self.kernel.K(x, self.Z), C, mvec)
        f_mean = mean + self.mean_function(x)
        f_var = (self.kernel.Kdiag(x) + q)[:, None].expand_as(f_mean)
        return f_mean, f_var

    def _init_posterior(self):
        """Initial q(u) from an exact GP on at most 100 random points (gptorch/models/sparse_gpr.py:310-335)."""
        i = np.random.permutation(self.Y.shape[0])[0: min(self.Y.shape[0], 100)]
        idx = torch.as_tensor(i, device=self.X.device)
        x, y = self.X[idx].to(self.compute_device), self.Y[idx].to(self.compute_device)
        likelihood = self.likelihood if isinstance(self.likelihood, Gaussian) else Gaussian(variance=0.01 * float(y.var()))
        model = GPR(x, y, self.kernel, mean_function=self.mean_function, likelihood=likelihood)
        with torch.no_grad():
            mean, cov = model._predict(self.Z.detach(), diag=False)
            mean = mean - self.mean_function(self.Z)
            chol_cov = cholesky(cov)
        return Param(mean.contiguous()), Param(chol_cov.contiguous(), transform=LowerCholeskyTransform())

    def _predict(self, x_new, diag=True, chol_kuu=None, _whitened=None, **kwargs):
        """q(f*) mean [n, dy] and variance [n, dy] / covariance [n, n] (gptorch/models/sparse_gpr.py:337-381)."""
        if chol_kuu is None and _whitened is None:
            def compute():
                L = cholesky(self.kernel.K(self.Z))
                return (L,) + tuple(self._whitened(L)[:2])
            chol_kuu, beta, t = self._memo("whitened", self.X, compute)
        else:
            chol_kuu = cholesky(self.kernel.K(self.Z)) if chol_kuu is None else chol_kuu
            beta, t = _whitened if _whitened is not None else self._whitened(chol_kuu)[:2]
        alpha = ag.TrsmRightFn.apply(self.kernel.K(x_new, self.Z), chol_kuu, ag._dinv_of(chol_kuu))   # [n, M]
        f_mean = mm(alpha, t) + self.mean_function(x_new)
        gamma = mm(alpha, beta, b_lower=True)      # beta = L^-1 L_S is lower triangular
        if diag:
            f_cov = (self.kernel.Kdiag(x_new) - ag.RowSumSqFn.apply(alpha) + ag.RowSumSqFn.apply(gamma))[:, None].expand_as(f_mean)
        else:
            f_cov = self.kernel.K(x_new) - mm_nt(alpha, alpha) + mm_nt(gamma, gamma)
        return f_mean, f_cov
