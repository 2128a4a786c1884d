"""Higher-precision witness for the GPR loss and its hyper-parameter gradients -- TEST INFRASTRUCTURE ONLY.

mpmath (40 significant digits) evaluation of what gptorch/models/gpr.py:47-67 defines mathematically, for a
stationary kernel on a SMALL data set (n ~ 100; everything is O(n^3) Python).  The scaled distance is formed from
coordinate differences, r_ij = sqrt(sum_d ((x_id - x_jd)/ell_d)^2), so r_ii = 0 exactly -- the reference's
|a|^2 + |b|^2 - 2 a.b form (gptorch/util.py:73-88) leaves O(1e-16) round-off there, which the Exp/Matern12
kernel turns into O(1e-8) through the square root (gptorch/kernels.py:171-172, 189-190).

Used by the tests to show where the documented Exp/Matern12 parity exception comes from: the CUDA result agrees with
this witness to the north-star tolerances (1e-9 / 1e-7) while the reference itself does not.
"""
import mpmath as mp

DPS = 40


def _k_and_dk(kind, r):
    """k(r)/sigma2 and the factor f(r) with d k / d log ell_d = sigma2 * f(r) * (delta_d/ell_d)^2."""
    if kind in ("Exp", "Matern12"):
        e = mp.e ** (-r)
        return e, (e / r if r != 0 else mp.mpf(0))
    if kind == "Rbf":
        e = mp.e ** (-r * r / 2)
        return e, e
    if kind == "Matern32":
        s = mp.sqrt(3) * r
        e = mp.e ** (-s)
        return (1 + s) * e, 3 * e
    if kind == "Matern52":
        s = mp.sqrt(5) * r
        e = mp.e ** (-s)
        return (1 + s + mp.mpf(5) / 3 * r * r) * e, mp.mpf(5) / 3 * (1 + s) * e
    raise ValueError(kind)


def gpr_loss_and_grads(kind, X, Y, ell, variance, noise):
    """(loss, {"variance", "length_scales", "noise"}) as Python floats: loss = -log p(Y) and its gradients with
    respect to the RAW (log) hyper-parameters, for dy = 1 and a zero mean function."""
    mp.mp.dps = DPS
    n, d = len(X), len(X[0])
    x = [[mp.mpf(float(v)) for v in row] for row in X]
    y = [mp.mpf(float(v[0])) for v in Y]
    ell = [mp.mpf(float(v)) for v in ell]
    s2, sn2 = mp.mpf(float(variance)), mp.mpf(float(noise))
    K = mp.matrix(n, n)
    dK = [mp.matrix(n, n) for _ in range(d)]
    for i in range(n):
        for j in range(i + 1):
            sq = [((x[i][k] - x[j][k]) / ell[k]) ** 2 for k in range(d)]
            r = mp.sqrt(mp.fsum(sq))
            kv, f = _k_and_dk(kind, r)
            K[i, j] = K[j, i] = s2 * kv
            for k in range(d):
                dK[k][i, j] = dK[k][j, i] = s2 * f * sq[k]
    Ky = K.copy()
    for i in range(n):
        Ky[i, i] += sn2
    L = mp.cholesky(Ky)
    Kinv = mp.inverse(Ky)
    yv = mp.matrix(y)
    a = Kinv * yv
    logdet = 2 * mp.fsum(mp.log(L[i, i]) for i in range(n))
    loss = (yv.T * a)[0] / 2 + logdet / 2 + mp.mpf(n) / 2 * mp.log(2 * mp.pi)
    W = Kinv - a * a.T

    def half_trace(A):      # 1/2 tr(W A), both symmetric
        return mp.fsum(W[i, j] * A[i, j] for i in range(n) for j in range(n)) / 2

    grads = {"variance": [float(half_trace(K))],
             "length_scales": [float(half_trace(dK[k])) for k in range(d)],
             "noise": [float(sn2 * mp.fsum(W[i, i] for i in range(n)) / 2)]}
    return float(loss), grads
