"""CPU tests of the oracle itself: known-answer vectors, base-R semantics, fixture integrity.
The reference's own tests pin no numbers for this path (tests/testthat/test-clone_id.R:7-49 are type
checks), so these pin the restatement against (a) Random123's published Philox vectors, (b) the fixtures
decoded from the reference's .RData files, (c) algebraic identities of the model."""
import math

import numpy as np
import pytest

from oracle import cardelino_oracle as O
from oracle import philox


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert [hex(int(x)) for x in philox.philox4x32(0, 0, 0, 0, 0, 0)] == \
        ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    f = 0xFFFFFFFF
    assert [hex(int(x)) for x in philox.philox4x32(f, f, f, f, f, f)] == \
        ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    assert [hex(int(x)) for x in philox.philox4x32(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344,
                                                    0xa4093822, 0x299f31d0)] == \
        ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]
    u = philox.uniforms(3, 1, 7, philox.STREAM_ASSIGN, np.arange(1000))
    assert u.min() > 0 and u.max() < 1 and abs(u.mean() - 0.5) < 0.05


def test_fixture_shapes(example_donor, simulation_input):
    A, D, Z = example_donor
    assert A.shape == (34, 428) and D.shape == (34, 428) and Z.shape == (34, 4)
    assert int(np.sum(np.nan_to_num(D) > 0)) == 1874
    assert np.nanmax(D) == 55 and np.nanmax(A) == 18 and np.nanmin(D) == 1
    assert np.all(Z[:, 0] == 0)
    D_input, Z4 = simulation_input
    assert D_input.shape == (439, 151) and Z4.shape == (403, 4)


def test_revsort_descending_and_sample_semantics():
    rng = np.random.default_rng(0)
    for _ in range(200):
        n = int(rng.integers(1, 9))
        a = list(rng.random(n).round(1))           # coarse values -> many ties
        ib = list(range(1, n + 1))
        a0 = list(a)
        O.revsort(a, ib)
        assert all(a[i] >= a[i + 1] for i in range(n - 1))
        assert sorted(ib) == list(range(1, n + 1))
        assert all(a0[ib[i] - 1] == a[i] for i in range(n))
    # sample(): first element of the descending CDF whose cumulative mass reaches u
    assert O.r_sample_one([0.1, 0.7, 0.2], 0.65) == 2
    assert O.r_sample_one([0.1, 0.7, 0.2], 0.75) == 3
    assert O.r_sample_one([0.1, 0.7, 0.2], 0.95) == 1
    assert O.r_sample_one([1.0], 0.3) == 1
    # frequencies
    u = np.random.default_rng(1).random(20000)
    lab = np.array([O.r_sample_one([0.2, 0.5, 0.3], x) for x in u])
    assert np.allclose(np.bincount(lab)[1:] / len(u), [0.2, 0.5, 0.3], atol=0.015)


def test_rbinom1_semantics():
    assert O.r_rbinom1(0.0, 0.999) == 0 and O.r_rbinom1(1.0, 0.001) == 1
    # p <= 0.5: 1 iff u >= 1 - p ; p > 0.5: 1 iff u < p
    assert O.r_rbinom1(0.3, 0.69) == 0 and O.r_rbinom1(0.3, 0.71) == 1
    assert O.r_rbinom1(0.8, 0.79) == 1 and O.r_rbinom1(0.8, 0.81) == 0
    u = np.random.default_rng(2).random(50000)
    assert abs(O.r_rbinom1(np.full_like(u, 0.37), u).mean() - 0.37) < 0.01


def test_preprocess_na_rules():
    A = np.array([[1.0, np.nan, 3.0], [np.nan, 2.0, 0.0]])
    D = np.array([[2.0, 4.0, 0.0], [np.nan, 2.0, 5.0]])
    A1, B1, W = O.preprocess(A, D)
    assert np.array_equal(A1, [[1, 0, 0], [0, 2, 0]])      # A NA & D>0 -> 0; D==0 -> missing
    assert np.array_equal(B1, [[1, 4, 0], [0, 0, 5]])
    assert W == pytest.approx(math.log(2) + 0 + 0 + 0)     # lchoose(2,1) + lchoose(4,0) + lchoose(2,2) + lchoose(5,0)


def test_em_known_answer(example_donor):
    """EM on example_donor converges from any start to the same optimum."""
    A, D, Z = example_donor
    for init in [(0.1, 0.5), (0.2, 0.35), (0.01, 0.59)]:
        r = O.clone_id_EM(A, D, Z, theta_init=init)
        assert r["theta"] == pytest.approx([0.03442, 0.66522], abs=2e-5)
        assert r["logLik"] == pytest.approx(-1854.0001, abs=2e-4)
        assert r["n_iter"] == 11
        assert list(np.bincount(r["prob"].argmax(1))) == [186, 197, 38, 7]
        assert np.allclose(r["prob"].sum(1), 1)


def test_loglik_dense_equals_table_identity(example_donor):
    A, D, Z = example_donor
    A1, B1, _ = O.preprocess(A, D)
    rng = np.random.default_rng(3)
    for _ in range(5):
        lab = rng.integers(1, 5, A.shape[1])
        cfg = (rng.random(Z.shape) < 0.4).astype(float)
        th1 = rng.uniform(0.05, 0.95, A.shape[0])
        dense = O.get_logLik(A1, B1, cfg, lab, 0.02, np.repeat(th1[:, None], A.shape[1], 1))
        table = O.get_logLik_tables(A1, B1, cfg, lab, 0.02, th1)
        assert dense == pytest.approx(table, rel=1e-12)


def test_geweke_matches_definition():
    rng = np.random.default_rng(4)
    X = rng.random((200, 7))
    Z = O.geweke_z(X)
    a, b = X[:20], X[99:]
    ref = (a.mean(0) - b.mean(0)) / np.sqrt(a.var(0, ddof=1) + ((b - b.mean(0)) ** 2).sum(0) / 19 + 1e-50)
    assert np.allclose(Z, ref)
    assert np.all(np.isnan(O.geweke_z(X[:5])))            # first window empty -> NaN like R


def test_geweke_from_segment_sums_is_the_same_test():
    """Streaming mode evaluates the stop rule from segment sums: the Z values agree with the
    definition to rounding, hence the same |Z| <= 2 decisions, on probability-like traces including saturated
    (constant 1 / 0) and tiny-valued columns, first row zero as in prob_all -- with fp64 and with fp32 segments."""
    rng = np.random.default_rng(11)
    n = 700
    X = np.column_stack([rng.random(n), 0.9 + 0.1 * rng.random(n), np.ones(n), np.zeros(n),
                         1e-30 * 10.0 ** (-8 * rng.random(n)), rng.beta(0.3, 0.3, n),
                         np.r_[rng.random(n // 2), 0.2 * rng.random(n - n // 2)]])       # last column: a chain that drifts
    X[0] = 0.0
    for rows in (n, 500, 300):
        Z = O.geweke_z(X[:rows])
        for storage, rtol in ((np.float64, 1e-6), (np.float32, 1e-3)):
            Zs = O.geweke_z_segment_sums(X[:rows], storage=storage)
            assert np.allclose(Z, Zs, rtol=rtol, atol=1e-9), (Z, Zs)
            assert np.array_equal(np.abs(Z) <= 2, np.abs(Zs) <= 2)


def test_segment_sums_survive_what_prefix_differences_lose():
    """Why the windows are sums of segments and not differences of running sums (round 1, advisor finding): columns
    that held an ordinary probability in a few early rows OUTSIDE both windows and sit at 1e-10 .. 1e-30 afterwards.
    The prefix-difference form loses window B's sums behind the early rows and flips verdicts; the segment form
    reproduces the two-pass definition."""
    rng = np.random.default_rng(0)
    n, cols = 20, 20000
    X = np.zeros((n, cols))
    hot = rng.random(cols) < 0.2
    for r in range(1, n):
        eps = 10.0 ** (-rng.uniform(3, 30, cols)) * (0.05 if r > 3 else 1.0)
        X[r] = np.where(hot, 1 - eps, eps)
        if r < 4:
            early = rng.random(cols) < 0.3
            X[r] = np.where(early, rng.random(cols), X[r])
    exact = np.abs(O.geweke_z(X)) <= 2
    seg = np.abs(O.geweke_z_segment_sums(X)) <= 2
    seg32 = np.abs(O.geweke_z_segment_sums(X, storage=np.float32)) <= 2
    pre = np.abs(O.geweke_z_prefix_differences(X)) <= 2
    assert (exact != seg).mean() == 0 and (exact != seg32).mean() < 1e-3
    assert (exact != pre).mean() > 5e-3                  # the old form flips a visible share of the verdicts


def test_deviance_ic():
    d = O.devianceIC(np.array([-10.0, -12.0, -11.0]), -9.0)
    assert d["DIC_Gelman"] == pytest.approx(18 + 4 * 1.0)
    assert d["DIC_Spiegelhalter"] == pytest.approx(18 + 2 * (22 - 18))


def test_colmatch():
    rng = np.random.default_rng(5)
    A = rng.random((30, 4))
    perm = np.array([2, 0, 3, 1])
    B = A[:, perm] + 0.01 * rng.random((30, 4))
    for force in (True, False):
        idx = O.colMatch(A, B, force=force)
        assert np.array_equal(perm[idx], np.arange(4))


def test_gibbs_dense_matches_stepwise_verifier(example_donor):
    """The free-running dense chain and the step-wise formulas agree (self-consistency of the oracle)."""
    A, D, Z = example_donor
    T = 25
    g = O.clone_id_Gibbs(A, D, Z, min_iter=T, max_iter=T, draws=O.NumpyDraws(11))
    assert g["prob_all"].shape == (T, 428 * 4) and np.all(g["prob_all"][0] == 0)
    assert np.allclose(g["prob_all"][1:].reshape(T - 1, 4, 428).sum(1), 1)
    # replay its own recorded draws through the verifier (uniforms are not recorded -> only prob/logLik)
    rec = dict(g)
    v = O.verify_gibbs_trace(A, D, Z, rec, min_iter=T, u_assign=np.full((T, 428), 2.0),
                             u_config=None, seed=0)
    assert v["max_dprob"] < 1e-12 and v["max_dloglik_rel"] < 1e-12


def test_gibbs_statistical_sanity(example_donor):
    """Order-of-magnitude expectations from the survey's probe (appendix A)."""
    A, D, Z = example_donor
    g = O.clone_id_Gibbs(A, D, Z, min_iter=300, max_iter=300, draws=O.NumpyDraws(7))
    assert 0.005 < g["theta0"] < 0.05
    assert 0.3 < np.nanmean(g["theta1"]) < 0.6
    assert -1350 < g["logLik"][-1] < -1100
    assert 0.5 < (g["prob"].max(1) > 0.5).mean() < 0.85


# ---- the sparse C restatement (oracle/gibbs_sparse.c) is pinned against the NumPy one ----

def _csc(A, D):
    A1, B1, W = O.preprocess(A, D)
    Dn = A1 + B1
    N, M = Dn.shape
    cols = [np.flatnonzero(Dn[:, j] > 0) for j in range(M)]
    p = np.concatenate([[0], np.cumsum([c.size for c in cols])]).astype(np.int64)
    vi = np.concatenate(cols).astype(np.int32)
    jj = np.repeat(np.arange(M), [c.size for c in cols])
    return A1, B1, W, p, vi, A1[vi, jj].astype(np.uint16), Dn[vi, jj].astype(np.uint16)


@pytest.mark.parametrize("wise", ["variant", "global"])
def test_c_oracle_sweep_matches_numpy_oracle(example_donor, c2_donor, wise):
    from oracle import c_oracle as CO
    rng = np.random.default_rng(4)
    for A, D, Z in (example_donor, c2_donor[:3]):
        A1, B1, W, p, vi, a, d = _csc(A, D)
        N, M = A1.shape
        K = Z.shape[1]
        cfg = (rng.random((N, K)) < 0.4).astype(float)
        cfg[:, 0] = 0
        th0 = 0.013
        th1 = rng.uniform(0.2, 0.8, N) if wise == "variant" else np.array([0.41])
        th1v = th1 if wise == "variant" else np.full(N, th1[0])
        Psi = rng.dirichlet(np.ones(K) * 5)
        u = rng.random(M)
        prob_o, lab_o = O.sweep_tables(A1, B1, cfg, th0, th1v, Psi, u)
        prob_c, z, SA, SB = CO.sweep_cells(N, M, K, p, vi, a, d, CO.config_masks(cfg), th0, th1, Psi, u)
        assert np.abs(prob_c - prob_o).max() <= 1e-12
        assert np.array_equal(z + 1, lab_o)                       # includes the all-tied uncovered cells (revsort)
        Zm = np.zeros((M, K)); Zm[np.arange(M), z] = 1
        assert np.array_equal(SA, (A1 @ Zm).astype(np.uint64)) and np.array_equal(SB, (B1 @ Zm).astype(np.uint64))
        # table step vs the dense formulas of R/clone_id.R:525-535,546-562 and get_logLik (:734-752)
        uc = rng.random(N * K)
        for relax_rate, keep_base in ((0.13, True), (0.5, False)):
            mask_new, S1, S2, S3, S4, ll, diff1 = CO.sweep_table(SA, SB, CO.config_masks(cfg), True, keep_base,
                                                                 relax_rate, th0, th1, uc, W)
            l0, l0c = np.log(th0), np.log(1 - th0)
            l1, l1c = np.log(th1v)[:, None], np.log(1 - th1v)[:, None]
            odd = SA * (l1 - l0) + SB * (l1c - l0c) + O._config_prior_oddlog(cfg, relax_rate, keep_base)
            odd = np.clip(odd, -50, 50)
            pflip = np.exp(odd) / (np.exp(odd) + 1)
            new = O.r_rbinom1(pflip.T.ravel(), uc).reshape(K, N).T
            assert np.array_equal(CO.masks_to_config(mask_new, K), new)
            assert S1 == (SA * (1 - new)).sum() and S2 == (SB * (1 - new)).sum()
            assert np.array_equal(S3, (SA * new).sum(1)) and np.array_equal(S4, (SB * new).sum(1))
            assert diff1 == int((new != cfg)[:, 1:].sum())
            ll_o = O.get_logLik(A1, B1, new, z + 1, th0, np.broadcast_to(th1v[:, None], A1.shape))
            assert abs(ll - ll_o) <= 1e-9 * abs(ll_o)


def test_c_oracle_runner_and_generator():
    """The timing runner recovers the labels of a donor made by the host generator (same distributions as the
    device generator): an end-to-end sanity check of the C restatement."""
    from oracle import c_oracle as CO
    import bench
    N, M, K, cov = 400, 3000, 5, 0.15
    cfg = bench.make_config(N, K, seed=2)
    p, vi, a, d, lab = CO.synth(N, M, K, cov, CO.config_masks(cfg), seed=9)
    assert p[-1] == vi.size and abs(p[-1] / (N * M) - cov) < 0.01 and np.all(a <= d) and d.min() >= 1
    assert all(np.all(np.diff(vi[p[j]:p[j + 1]]) > 0) for j in range(0, M, 97))
    sec, z, th1 = CO.run(N, M, K, p, vi, a, d, CO.config_masks(cfg), 30, seed=1)
    assert sec > 0 and (z + 1 == lab).mean() > 0.95 and np.all((th1 > 0) & (th1 < 1))


def test_oracle_reproduces_committed_golden_vectors(example_donor):
    """tests/golden/oracle_example_donor.npz (tools/make_golden.py): replaying the recorded uniforms and Beta draws
    reproduces the recorded chain; EM from the stored start and get_logLik at the stored state give the stored
    numbers.  Pins the oracle against drift (the reference itself cannot run here: parity unpinned)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_example_donor.npz"))
    A, D, Z = example_donor
    T = int(g["T"])
    draws = O.ReplayDraws(theta0_all=g["theta0_all"], theta1_all=g["theta1_all"], relax_rate_all=g["relax_rate_all"],
                          u_assign=g["u_assign"], u_config=g["u_config"])
    o = O.clone_id_Gibbs(A, D, Z, min_iter=T, max_iter=T, draws=draws)
    assert np.array_equal(o["assign_all"], g["assign_all"]) and np.array_equal(o["Config_all"], g["Config_all"])
    assert np.abs(o["prob"] - g["prob"]).max() <= 1e-12 and np.abs(o["prob_all"][T - 1] - g["prob_all_last"]).max() <= 1e-12
    assert np.allclose(o["logLik"], g["logLik"], rtol=1e-12, atol=0)
    assert np.array_equal(o["Config_prob"], g["Config_prob"])
    em = O.clone_id_EM(A, D, Z, theta_init=g["em_theta_init"])
    assert em["n_iter"] == int(g["em_n_iter"]) and np.allclose(em["theta"], g["em_theta"], rtol=1e-12)
    assert em["logLik"] == pytest.approx(float(g["em_logLik"]), rel=1e-12)
    A1, B1, _ = O.preprocess(A, D)
    it, K, N = int(g["ll_it"]), Z.shape[1], Z.shape[0]
    ll = O.get_logLik_tables(A1, B1, g["Config_all"][it - 1].reshape(K, N).T.astype(float), g["assign_all"][it - 1],
                             g["theta0_all"][it - 1], g["theta1_all"][it - 1])
    assert ll == pytest.approx(float(g["ll_value"]), rel=1e-12)
