"""The arithmetic of the experimental int8-sliced engine (csrc/gpb_ozaki.cu), restated in numpy on the CPU tier: the split
is error-free, every slice fits the int8 range the int32 accumulation bound assumes, one weight class is ONE product over the
concatenated k (forward slices of A against reversed slices of B), and the truncated sum converges at 7 bits per slice."""
import numpy as np

BITS = 7


def split(A, S):
    amax = np.abs(A).max(axis=1, keepdims=True)
    e = np.where(amax > 0, np.floor(np.log2(np.where(amax > 0, amax, 1.0))) + 2.0, 0.0)      # |A| 2^-e < 1/2 (ilogb + 2)
    r = A * np.exp2(-e)
    qs = []
    for _ in range(S):
        r = r * 2.0 ** BITS
        q = np.rint(r)
        r = r - q
        qs.append(q.astype(np.int8))
    return qs, np.exp2(e), r


def sliced_matmul_nt(A, B, S):
    qa, sa, _ = split(A, S)
    qb, sb, _ = split(B, S)
    k = A.shape[1]
    a8 = np.concatenate(qa, axis=1).astype(np.int32)                        # [Q0 | Q1 | ... ]
    b8 = np.concatenate(qb[::-1], axis=1).astype(np.int32)                  # [Q_{S-1} | ... | Q0]
    t = np.zeros((A.shape[0], B.shape[0]))
    for u in range(S - 1, -1, -1):                                          # Horner from the smallest weight class up
        p = a8[:, : (u + 1) * k] @ b8[:, (S - 1 - u) * k:].T                # prefix of A against suffix of B
        assert np.abs(p).max() < 2 ** 31
        pairwise = sum(qa[s].astype(np.int64) @ qb[u - s].astype(np.int64).T for s in range(u + 1))
        assert np.array_equal(p, pairwise)                                  # one GEMM per weight class == the pairwise sum
        t = t * 2.0 ** -BITS + p
    return (sa * sb.T) * (t * 2.0 ** (-2 * BITS))


def test_split_is_error_free_and_slices_fit_seven_bits():
    rng = np.random.default_rng(0)
    A = rng.standard_normal((40, 96)) * np.exp(6 * rng.standard_normal((40, 96)))
    A[3] = 0.0                                                              # an all-zero row
    A[5, 0] = A[5].max() * 2.0 ** 40                                        # one dominant entry
    S = 9
    qs, scale, rem = split(A, S)
    assert max(int(np.abs(q.astype(np.int32)).max()) for q in qs) <= 64
    rec = sum(q.astype(np.float64) * 2.0 ** (-BITS * (s + 1)) for s, q in enumerate(qs))
    assert np.array_equal((rec + rem * 2.0 ** (-BITS * S)) * scale, A)      # exact: remainder included
    assert np.abs(rec * scale - A).max() <= 2.0 ** (-BITS * S - 1) * scale.max() * 1.0000001


def test_weight_classes_as_single_products_and_convergence():
    rng = np.random.default_rng(1)
    A = rng.standard_normal((48, 64)) * np.exp(2 * rng.standard_normal((48, 1)))
    B = rng.standard_normal((32, 64))
    ref = (A.astype(np.longdouble) @ B.astype(np.longdouble).T).astype(np.float64)
    scale = np.abs(A).max(axis=1, keepdims=True) * np.abs(B).max(axis=1, keepdims=True).T * A.shape[1]
    errs = {S: float((np.abs(sliced_matmul_nt(A, B, S) - ref) / scale).max()) for S in (2, 4, 6, 8)}
    for S, e in errs.items():
        assert e <= 2.0 ** (-BITS * S + 2), (S, e)                          # the documented norm-wise bound
    assert errs[8] < 1e-15
