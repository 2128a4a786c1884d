"""Utilities with the reference's names (gptorch/util.py)."""
import numpy as np
import torch
from scipy.cluster.vq import kmeans2

from . import settings

TensorType = torch.DoubleTensor   # gptorch/util.py:11 -- kept for isinstance/constructor compatibility
torch_dtype = torch.double        # gptorch/util.py:12


def as_tensor(x, device=None):
    """numpy array / float / tensor -> float64 tensor on the package's default device (gptorch/util.py:15-31)."""
    device = settings.default_device() if device is None else device
    if isinstance(x, torch.Tensor):
        return x.to(dtype=torch_dtype, device=device)
    if isinstance(x, np.ndarray):
        return torch.as_tensor(x, dtype=torch_dtype).to(device)
    if isinstance(x, float):
        return torch.tensor([x], dtype=torch_dtype, device=device)
    raise TypeError("Unsupported type {}".format(type(x)))


def kmeans_centers(x, k, perturb_if_fail=False):
    """Cluster centres used to initialise inducing inputs (gptorch/util.py:34-49); host-side, run once."""
    try:
        return kmeans2(x, k)[0]
    except np.linalg.LinAlgError:
        if not perturb_if_fail:
            raise
        jiggle = 1.0e-4 * x.std(axis=0) * np.random.randn(*x.shape)
        return kmeans2(x + jiggle, k)[0]


KMEANS_HOST_MAX_ROWS = 200000     # above this, scipy's host k-means takes minutes: cluster on the device
KMEANS_DEVICE_SAMPLE = 1 << 22    # rows of a larger data set used for the device k-means (uniform subsample)


def kmeans_centers_device(x, k, iters=10, chunk=1 << 20):
    """Lloyd's k-means on the device for data sets where the reference's host scipy.cluster.vq.kmeans2
    (gptorch/util.py:34-49, 10 iterations) would take minutes.  Initial centres are k distinct rows
    drawn with the host numpy RNG (kmeans2's minit="points"); each iteration assigns rows by the largest
    x.c - |c|^2/2 (the x.c^T product on the native FP64 engine, in row chunks) and moves every non-empty centre to the
    mean of its rows.  Returns a numpy array [k, D] like kmeans_centers."""
    from . import _native as nv
    n = x.shape[0]
    dev = settings.default_device()
    if n > KMEANS_DEVICE_SAMPLE:
        rows = np.sort(np.random.default_rng(np.random.randint(1 << 31)).choice(n, size=KMEANS_DEVICE_SAMPLE, replace=False))
        x = x[torch.from_numpy(rows)] if isinstance(x, torch.Tensor) else x[rows]
        n = KMEANS_DEVICE_SAMPLE
    X = as_tensor(x, device=dev)
    C = X[torch.as_tensor(np.random.choice(n, size=k, replace=False), device=dev)].clone()
    d = X.shape[1]
    for _ in range(iters):
        sums = torch.zeros((k, d), dtype=torch_dtype, device=dev)
        counts = torch.zeros(k, dtype=torch_dtype, device=dev)
        half_norm = 0.5 * (C * C).sum(1)
        for s in range(0, n, chunk):
            Xc = X[s: s + chunk]
            score = nv.gemm(nv.GEMM_NT, Xc, C) - half_norm          # argmax of x.c - |c|^2/2 == argmin of |x - c|^2
            a = score.argmax(1)
            sums.index_add_(0, a, Xc)
            counts += torch.bincount(a, minlength=k).to(torch_dtype)
        C = torch.where((counts > 0)[:, None], sums / counts.clamp(min=1.0)[:, None], C)
    return C.cpu().numpy()


def PCA(X, q):
    """Project X (n, p) on its q leading principal directions (gptorch/util.py:52-70)."""
    assert q <= X.shape[1], "Cannot have more latent dimensions than observed"
    evals, evecs = np.linalg.eigh(np.cov(X.T))
    order = np.argsort(evals)[::-1][:q]
    return (X - X.mean(0)).dot(evecs[:, order])


def squared_distance(x1, x2=None):
    """Pairwise squared distances, [n1, n2] (gptorch/util.py:73-88).

    Kept as composite torch ops on purpose: the reference's tests differentiate this function twice
    (test/test_util.py:80-106) and the detach-clamp must leave both derivatives intact.  Kernel.K does not go
    through here; it calls the fused CUDA kernel.
    """
    if x2 is None:
        x2 = x1
    sq1 = (x1 * x1).sum(1, keepdim=True)
    sq2 = (x2 * x2).sum(1, keepdim=True)
    r2 = sq1 + sq2.t() - 2.0 * x1 @ x2.t()
    return r2 - torch.clamp(r2, max=0.0).detach()
