"""The SCHEDULE of the pipelined factorisation (csrc/kernels_chol_pipe.cu + k_side_updates), restated on the CPU job by job
(oracle/pipe_oracle.py), against numpy.linalg: every tile must receive every update exactly once, nothing may be read before its first
writer (the working buffers start as NaN -- the CUDA path has no memsets either), nothing may be enqueued before what it depends on,
and the jobs of one launch must be independent of each other (launch A is executed in reverse job order)."""
import numpy as np
import pytest

from oracle import pipe_oracle as P


def _spd(n, seed):
    rng = np.random.default_rng(seed)
    M = rng.standard_normal((n, n))
    return M @ M.T / n + np.eye(n)


@pytest.mark.parametrize("T", [2, 3, 4, 5, 8, 9, 12, 13, 16])
@pytest.mark.parametrize("want_kinv", [True, False])
def test_pipelined_schedule_matches_numpy(T, want_kinv):
    b = 4
    A = _spd(T * b, T)
    st = P.chol_inverse_pipelined(A, b, want_kinv=want_kinv, cap=6, sms=8)
    L = np.linalg.cholesky(A)
    Li = np.linalg.inv(L)
    # the factor: diagonal tiles + every panel tile
    assert np.allclose(np.tril(st.L), L, rtol=0, atol=1e-12 * np.abs(L).max())
    assert np.allclose(st.Li, Li, rtol=0, atol=1e-11 * np.abs(Li).max()) and not np.isnan(st.Li).any()
    assert np.allclose(st.Lt, Li.T, rtol=0, atol=1e-11 * np.abs(Li).max())
    if want_kinv:
        Kinv = Li.T @ Li
        got = st.Kinv.copy()
        for a in range(T):                                           # lower TILES are produced (diagonal tiles whole)
            for c in range(a + 1):
                t = got[a * b:(a + 1) * b, c * b:(c + 1) * b]
                assert not np.isnan(t).any(), (a, c)
                assert np.allclose(t, Kinv[a * b:(a + 1) * b, c * b:(c + 1) * b], rtol=0, atol=1e-10 * np.abs(Kinv).max()), (a, c)


@pytest.mark.parametrize("cap,sms", [(1, 1), (3, 4), (7, 9), (50, 60), (1000, 1000)])
def test_result_does_not_depend_on_the_grid_cap(cap, sms):
    T, b = 9, 4
    A = _spd(T * b, 5)
    ref = P.chol_inverse_pipelined(A, b, True, cap=6, sms=8)
    st = P.chol_inverse_pipelined(A, b, True, cap=cap, sms=sms)
    # the cut only moves the launch boundary: the same jobs on the same tiles in the same order per tile -> identical bits
    assert np.array_equal(st.Li, ref.Li)
    assert np.array_equal(np.nan_to_num(st.Kinv), np.nan_to_num(ref.Kinv))
    for p, cut, total, n_head in st.launches:
        assert 0 <= cut <= total


def test_jobs_of_one_launch_are_independent():
    """Launch A always runs in reverse job order in the oracle; here launch B too: the per-tile update order is unchanged (a tile has
    one job per launch), so the results must be bit-identical."""
    T, b = 13, 4
    A = _spd(T * b, 7)
    fwd = P.chol_inverse_pipelined(A, b, True, cap=6, sms=8)
    rev = P.chol_inverse_pipelined(A, b, True, cap=6, sms=8, reverse_b=True)
    assert np.array_equal(fwd.Li, rev.Li) and np.array_equal(np.nan_to_num(fwd.Kinv), np.nan_to_num(rev.Kinv))
    assert np.array_equal(np.nan_to_num(fwd.L), np.nan_to_num(rev.L))


def test_job_counts_per_pair_are_uniform_at_the_headline_size():
    """T = 32 (n = 4096): every middle pair has 522 jobs -- the figure DESIGN.md 2c quotes for the 4-rounds-per-pair bound."""
    T = 32
    counts = []
    p_last = (T + 1) // 2 - 1
    for p in range(p_last + 1):
        c0, c1 = 2 * p, 2 * p + 1
        n0 = c1 + 1
        nn = min(2, max(T - n0, 0))
        mb = max(T - n0 - 2, 0)
        n = mb * nn + 2 * c0 + mb * (mb + 1) // 2 + (c1 + 1) * nn + (c1 + 1) * mb
        n += c0 * (c0 + 1) // 2 if 1 <= p < p_last - 1 else 0
        n += (c1 + 1) * (c1 + 2) // 2 if p >= p_last - 1 else 0
        counts.append(n)
    assert set(counts[1:p_last - 1]) == {522}


def test_lower_tile_decode_roundtrip():
    t = 0
    for bm in range(40):
        for bn in range(bm + 1):
            assert P.lower_tile_decode(t) == (bm, bn)
            t += 1
