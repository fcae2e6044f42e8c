"""CPU restatement (numpy, float64) of the PIPELINED tile factorisation of hdbo_benchmark_b200/csrc/kernels_chol_pipe.cu +
k_side_updates (csrc/kernels_gemm.cu): right-looking Cholesky + inverse + K^-1 in pairs of tile columns.

TEST INFRASTRUCTURE ONLY (like everything under oracle/): imported by tests/, never by the package.

What it restates is not upstream arithmetic (GPyTorch calls LAPACK potrf / potri here) but the SCHEDULE of the CUDA path, launch by
launch and job by job, in the order the host enqueues it -- the same regions, the same index decode of the job list (head-first order,
the cut after one round of the grid), the same "first job on a tile starts from zero" rule that replaces the memsets, the K^-1 jobs
run one pair late and merged into K = 512 jobs for the last pair but one, panels and rows of L^-1 through the pair's 2x2-tile
inverse.  Working buffers start as NaN: a tile that is read before its first writer, an update applied twice or never, or an
operation enqueued before what it depends on shows up as a wrong / NaN result against numpy.linalg.
Each function cites the CUDA code it follows (file: function).
"""
from __future__ import annotations

import numpy as np


def lower_tile_decode(t: int):
    """kernels_gemm.cu: lower_tile_decode -- t enumerates (bm, bn), bm >= bn, row by row: t = bm (bm + 1) / 2 + bn."""
    r = int((np.sqrt(8.0 * t + 1.0) - 1.0) * 0.5)
    while (r + 1) * (r + 2) // 2 <= t:
        r += 1
    while r * (r + 1) // 2 > t:
        r -= 1
    return r, t - r * (r + 1) // 2


class PipeState:
    """The buffers of hdbo_gp the schedule touches, as tile-addressed float64 arrays (tile edge b)."""

    def __init__(self, A: np.ndarray, b: int, want_kinv: bool):
        n = A.shape[0]
        assert n % b == 0
        self.b, self.T = b, n // b
        nan = np.full((n, n), np.nan)
        self.Aw = nan.copy()                      # lower tiles: Khat (consumed); upper tiles: W^T, NEVER cleared (first-touch rule)
        il = np.tril_indices(n)
        self.Aw[il] = A[il]
        for t in range(self.T):                   # the build writes whole diagonal tiles
            s = slice(t * b, (t + 1) * b)
            self.Aw[s, s] = A[s, s]
        self.L = nan.copy()
        self.Li = np.zeros((n, n))                # engine.py allocates L^-1 / L^-T with zeros: their zero halves are relied upon
        self.Lt = np.zeros((n, n))
        self.Kinv = nan.copy() if want_kinv else None
        self.launches = []                        # (name, jobs) log for the tests

    def tile(self, M, r, c, h=1, w=1):
        b = self.b
        return M[r * b:(r + h) * b, c * b:(c + w) * b]


def leaf(st: PipeState, k: int):
    """leaf.cuh: k_leaf_blocked -- Cholesky of the diagonal tile (lower part read) and its triangular inverse (+ transpose)."""
    a = np.tril(st.tile(st.Aw, k, k))
    a = a + np.tril(a, -1).T
    l = np.linalg.cholesky(a)
    li = np.linalg.inv(l)
    st.tile(st.L, k, k)[:] = l
    st.tile(st.Li, k, k)[:] = np.tril(li)
    st.tile(st.Lt, k, k)[:] = np.tril(li).T


def panel(st: PipeState, k: int, w: int, r0: int, nr: int):
    """kernels_chol_pipe.cu: panel -- L(rows, k..k+w-1) = A(rows, k..) Linv(k..k+w-1, k..k+w-1)^T (KR_LE_N: lower-triangular B)."""
    if nr <= 0:
        return
    st.tile(st.L, r0, k, nr, w)[:] = st.tile(st.Aw, r0, k, nr, w) @ np.tril(st.tile(st.Li, k, k, w, w)).T


def update_col(st: PipeState, c: int, r0: int, nr: int, k0: int, w: int):
    """kernels_chol_pipe.cu: update_col -- A(rows, c) -= L(rows, k0..) L(c, k0..)^T."""
    if nr <= 0:
        return
    st.tile(st.Aw, r0, c, nr, 1)[:] -= st.tile(st.L, r0, k0, nr, w) @ st.tile(st.L, c, k0, 1, w).T


def side_job(st: PipeState, sp: dict, job: int):
    """kernels_gemm.cu: k_side_updates -- one job of the pair's list (regions a, s, c, e, b, d, e2 in list order)."""
    T, b = st.T, st.b
    k0, w, n0, nn, f, jmax = sp["k0"], sp["w"], sp["n0"], sp["nn"], sp["f"], sp["jmax"]
    mb = T - f
    e_a = sp["n_a"]; e_s = e_a + sp["n_s"]; e_c = e_s + sp["n_c"]; e_e = e_c + sp["n_e"]; e_b = e_e + sp["n_b"]; e_d = e_b + sp["n_d"]
    kw = w                                                       # tile columns of the contraction

    def acc(C, first, sign, A, B):
        if first:
            C[:] = sign * (A @ B.T)
        else:
            C[:] += sign * (A @ B.T)

    if job < e_a:
        j, i = n0 + job % nn, f + job // nn
        acc(st.tile(st.Aw, i, j), False, -1.0, st.tile(st.L, i, k0, 1, kw), st.tile(st.L, j, k0, 1, kw))
    elif job < e_s:
        q = job - e_a
        r, j = q % w, q // w
        res = st.tile(st.Li, k0 + r, k0, 1, r + 1) @ st.tile(st.Aw, j, k0, 1, r + 1).T     # Linv(k0+r, k0..k0+r) W(k0..k0+r, j)
        st.tile(st.Li, k0 + r, j)[:] = res
        st.tile(st.Lt, j, k0 + r)[:] = res.T
    elif job < e_c:
        bi, bj = lower_tile_decode(sp["n_c"] - 1 - (job - e_s))
        i, j = f + (mb - 1 - bj), f + (mb - 1 - bi)
        acc(st.tile(st.Aw, i, j), False, -1.0, st.tile(st.L, i, k0, 1, kw), st.tile(st.L, j, k0, 1, kw))
    elif job < e_e:
        a, bb = lower_tile_decode(job - e_c)
        ek0, ekw = sp["e_k0"], 2
        acc(st.tile(st.Kinv, a, bb), a >= ek0, 1.0, st.tile(st.Lt, a, ek0, 1, ekw), st.tile(st.Lt, bb, ek0, 1, ekw))
    elif job < e_b:
        q = job - e_e
        j, i = q % (jmax + 1), n0 + q // (jmax + 1)
        acc(st.tile(st.Aw, j, i), j >= k0, -1.0, st.tile(st.Lt, j, k0, 1, kw), st.tile(st.L, i, k0, 1, kw))
    elif job < e_d:
        q = job - e_b
        j, i = q % (jmax + 1), f + q // (jmax + 1)
        acc(st.tile(st.Aw, j, i), j >= k0, -1.0, st.tile(st.Lt, j, k0, 1, kw), st.tile(st.L, i, k0, 1, kw))
    else:
        a, bb = lower_tile_decode(job - e_d)
        ek0, ekw = sp["e2_k0"], sp["e2_kw"]
        acc(st.tile(st.Kinv, a, bb), a >= ek0, 1.0, st.tile(st.Lt, a, ek0, 1, ekw), st.tile(st.Lt, bb, ek0, 1, ekw))


def chol_inverse_pipelined(A: np.ndarray, b: int, want_kinv: bool = True, cap: int = 6, sms: int = 8, reverse_b: bool = False) -> PipeState:
    """kernels_chol_pipe.cu: chol_inverse_pipelined -- every launch in host enqueue order (a sequential execution in that order is a
    valid execution only if nothing is enqueued before its dependencies: that is part of what the tests check).
    `cap` / `sms` play the roles of the side-grid cap (140) and the SM count (148); `reverse_b` runs the jobs of launch B backwards too."""
    st = PipeState(A, b, want_kinv)
    T = st.T
    p_last = (T + 1) // 2 - 1
    p = 0
    for c0 in range(0, T, 2):
        c1 = c0 + 1
        has_c1 = c1 < T
        last = c1 if has_c1 else c0
        w = 2 if has_c1 else 1
        n0 = last + 1
        nn = min(2, max(T - n0, 0))
        f = n0 + 2
        mb = max(T - f, 0)
        leaf(st, c0)
        if has_c1:
            panel(st, c0, 1, c1, 1)                               # chain: L(c1, c0)
            update_col(st, c1, c1, 1, c0, 1)                      # chain: A(c1, c1)
            leaf(st, c1)
            if nn > 0:                                            # panel stream, beside leaf(c1)
                panel(st, c0, 1, n0, nn)
                update_col(st, c1, n0, nn, c0, 1)
            # near stream: W^T(c0, c1) = -L00^-T L(c1, c0)^T (KR_GE_M: upper-triangular A), then Linv(c1, c0) = L11^-1 W(c1, c0)
            st.tile(st.Aw, c0, c1)[:] = -(np.triu(st.tile(st.Lt, c0, c0)) @ st.tile(st.L, c1, c0).T)
            r = np.tril(st.tile(st.Li, c1, c1)) @ st.tile(st.Aw, c0, c1).T
            st.tile(st.Li, c1, c0)[:] = r
            st.tile(st.Lt, c0, c1)[:] = r.T
        if nn > 0:
            panel(st, c1, 1, n0, nn)                              # chain: L(n0..n1, c1)
            # chain: the next pair's diagonal block, lower tiles (lower_only GEMM, K = 256)
            for i in range(nn):
                for j in range(i + 1):
                    st.tile(st.Aw, n0 + i, n0 + j)[:] -= st.tile(st.L, n0 + i, c0, 1, w) @ st.tile(st.L, n0 + j, c0, 1, w).T
            if mb > 0:                                            # panel stream: the rest of the pair's panel through the pair inverse
                panel(st, c0, w, f, mb)
        # ---- side: the job list of this pair, launch A = jobs [0, cut), launch B = the rest
        kinv_on = want_kinv
        sp = dict(k0=c0, w=w, n0=n0, nn=nn, f=f, jmax=last, e_k0=c0 - 2)
        sp["n_a"] = mb * nn
        sp["n_s"] = w * c0
        sp["n_c"] = mb * (mb + 1) // 2
        sp["n_e"] = c0 * (c0 + 1) // 2 if (kinv_on and 1 <= p < p_last - 1) else 0
        sp["n_b"] = (last + 1) * nn
        sp["n_d"] = (last + 1) * mb
        sp["n_e2"] = (last + 1) * (last + 2) // 2 if (kinv_on and p >= p_last - 1) else 0
        merged = (p == p_last - 1) and p >= 1
        sp["e2_k0"] = c0 - 2 if merged else c0
        sp["e2_kw"] = w + 2 if merged else w
        n_l = sp["n_a"] + sp["n_s"] + sp["n_c"] + sp["n_e"]
        total = n_l + sp["n_b"] + sp["n_d"] + sp["n_e2"]
        n_head = sp["n_a"] + sp["n_s"] + min(sp["n_c"], mb + 1)
        ctas = cap if nn > 0 else sms
        cut = min(max(n_head, ctas), n_l)
        # a launch's jobs run concurrently on the GPU: inside one launch no job may read what another job of it writes.  Executing
        # launch A in REVERSE job order and launch B in forward order would expose such a dependence as an order sensitivity.
        for job in reversed(range(0, cut)):
            side_job(st, sp, job)
        for job in (reversed(range(cut, total)) if reverse_b else range(cut, total)):
            side_job(st, sp, job)
        st.launches.append((p, cut, total, n_head))
        if nn == 0:
            break
        p += 1
    return st
