"""Boundary-MPS restatement (dense, bosonic).  Test infrastructure (see oracle/__init__.py).

Follows one_dim_tn/boundary_mps/bmps_impl.h of the reference:
  * vacuum BMPS                 bmps_impl.h:21-96
  * RightCanonicalizeTruncate   bmps_impl.h:225-263
  * ReverseTransferMPOIfNeeded_ bmps_impl.h:694-699
  * MultiplyMPOSVDCompress_     bmps_impl.h:756-862
TensorToolkit ``Contract``/``QR``/``SVD`` (absent from the container) are restated as
``numpy.einsum`` / LAPACK geqrf+orgqr / gesdd; the truncation rule is TensorToolkit's
"smallest t in [Dmin, Dmax] with discarded weight <= trunc_err" (SURVEY.md section 8a row 7).

Conventions (identical to the reference): positions LEFT=0, DOWN=1, RIGHT=2, UP=3
(basic.h:59-64); site tensor legs (L, D, R, U); BMPS tensor legs (0, phys, 2) in STORAGE order,
UP/RIGHT stored reversed (bmps.h:146-152).
"""
import numpy as np


def es(spec, *ops):
    """einsum routed through BLAS (tensordot) -- what TensorToolkit's Contract does via GEMM."""
    return np.einsum(spec, *ops, optimize=True)


LEFT, DOWN, RIGHT, UP = 0, 1, 2, 3
HORIZONTAL, VERTICAL = 0, 1


def opposite(post):
    return (post + 2) % 4


def mpo_perm(post):
    """Leg permutation (pre_post, post, next_post, opposite) of a site tensor seen from a BMPS/BTen at
    ``post`` (bmps.h:273-277: PrePostLegIndex_ = (post+3)%4, NextPostLegIndex_ = (post+1)%4)."""
    return ((post + 3) % 4, post, (post + 1) % 4, (post + 2) % 4)


def truncation_dim(s, dmin, dmax, trunc_err):
    """Number of kept singular values: TensorToolkit SVD(..., trunc_err, Dmin, Dmax) semantics."""
    n = len(s)
    if n <= dmin:
        return n
    total = float(np.sum(s * s))
    kept = n
    kept_sum = total
    while kept > dmin:
        sv2 = float(s[kept - 1] ** 2)
        if kept <= dmax and total > 0 and (1.0 - (kept_sum - sv2) / total) > trunc_err:
            break
        kept_sum -= sv2
        kept -= 1
    return kept


def vacuum_bmps(n):
    """BMPS(position, hilbert_spaces) with all legs of dimension 1 (bmps_impl.h:59-96)."""
    return [np.ones((1, 1, 1)) for _ in range(n)]


def multiply_mpo(mps, mpo_sites, post, dmin, dmax, trunc_err, stats=None):
    """BMPS::MultiplyMPO with SVD_COMPRESS (bmps_impl.h:404-437 -> 756-862).

    mps        list of (a, p, b) tensors in storage order
    mpo_sites  list of site tensors (L, D, R, U) in LATTICE order along the slice
               (row: increasing col; column: increasing row) -- reversed here for RIGHT/UP
               exactly like ReverseTransferMPOIfNeeded_.
    """
    n = len(mps)
    assert len(mpo_sites) == n
    mpo = list(mpo_sites)
    if post in (RIGHT, UP):
        mpo = mpo[::-1]
    perm = mpo_perm(post)
    dtype = np.result_type(mps[0].dtype, mpo[0].dtype)
    res = [None] * n
    r = np.ones((1, 1, 1), dtype=dtype)          # r[k, e(mpo), a(mps)]   bmps_impl.h:767-774
    for i in range(n):
        m = np.transpose(mpo[i], perm)            # m[e, p, f, o]
        tmp1 = es("apb,kea->pbke", mps[i], r)        # bmps_impl.h:806
        tmp2 = es("pbke,epfo->bkfo", tmp1, m)        # bmps_impl.h:807
        if i < n - 1:
            t = np.transpose(tmp2, (1, 3, 2, 0))            # (k, o, f, b)  bmps_impl.h:811-817
            k, o, f, b = t.shape
            q, rr = np.linalg.qr(t.reshape(k * o, f * b), mode="reduced")   # bmps_impl.h:821
            j = q.shape[1]
            res[i] = q.reshape(k, o, j)
            r = rr.reshape(j, f, b)
        else:
            assert tmp2.shape[0] == 1 and tmp2.shape[2] == 1
            res[i] = np.ascontiguousarray(tmp2[0, :, 0, :])[:, :, None]   # (k, o, 1)  bmps_impl.h:826-838
    dmax_seen, err_seen = 1, 0.0
    for i in range(n - 1, 0, -1):                 # bmps_impl.h:853-857
        d_i, e_i = right_canonicalize_truncate(res, i, dmin, dmax, trunc_err)
        dmax_seen, err_seen = max(dmax_seen, d_i), max(err_seen, e_i)
    if stats is not None:
        stats["D"] = dmax_seen
        stats["trunc_err"] = err_seen
    return res


def right_canonicalize_truncate(res, site, dmin, dmax, trunc_err):
    """BMPS::RightCanonicalizeTruncate (bmps_impl.h:225-263)."""
    a = res[site]
    k, o, j = a.shape
    u, s, vt = np.linalg.svd(a.reshape(k, o * j), full_matrices=False)
    t = truncation_dim(s, dmin, dmax, trunc_err)
    total = float(np.sum(s * s))
    err = float(np.sum(s[t:] ** 2) / total) if total > 0 else 0.0
    res[site] = vt[:t].reshape(t, o, j)
    us = u[:, :t] * s[:t]
    res[site - 1] = es("aok,kt->aot", res[site - 1], us)
    return t, err
