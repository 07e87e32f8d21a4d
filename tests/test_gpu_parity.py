"""GPU parity tests (run with -m gpu on a B200).  Every test calls the CUDA path through the C ABI
(ctypes -> libcardelino_b200.so) and checks it against the NumPy oracle on the same inputs.

Tolerances (BASELINE.json north_star): get_logLik / EM <= 1e-6 relative; replayed chains reproduce
assignments exactly except draws within 1e-6 of a CDF boundary (or whose order in sample()'s sort is
decided by rounding between two probabilities equal to 1e-12); converged prob within Monte-Carlo error
(max |dprob| <= 0.02 on decided cells, identical argmax where prob > 0.9)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from oracle import cardelino_oracle as O  # noqa: E402

REL = 1e-6


def _cb():
    import cardelino_b200 as cb
    return cb


def _verify(cb, h, A, D, cfg, T=40, seed=11, **kw):
    res = cb.clone_id_Gibbs(A, D, cfg, min_iter=T, max_iter=T, seed=seed, keep_assign_all=True, handle=h,
                            verbose=False, **kw)
    okw = {k: v for k, v in kw.items() if k in ("relax_Config", "relax_rate_fixed", "wise", "Psi",
                                                "keep_base_clone", "relax_rate_prior")}
    v = O.verify_gibbs_trace(A, D, cfg, res, min_iter=T, seed=seed, **okw)
    assert v["max_dprob"] <= 1e-9, v
    assert v["hard"] == 0 and v["config_hard"] == 0, v
    assert v["max_dloglik_rel"] <= REL, v
    n_lab = (T - 1) * np.asarray(A).shape[1]
    assert v["boundary"] + v["near_tie"] <= max(2, 0.02 * n_lab), v
    return res, v


def test_em_example_donor(gpu_handle, example_donor):
    cb = _cb()
    A, D, Z = example_donor
    for init in [(0.1, 0.5), (0.2, 0.35)]:
        g = cb.clone_id_EM(A, D, Z, theta_init=init, handle=gpu_handle, verbose=False)
        o = O.clone_id_EM(A, D, Z, theta_init=init)
        assert g["n_iter"] == o["n_iter"]
        assert np.allclose(g["theta"], o["theta"], rtol=REL, atol=0)
        assert abs(g["logLik"] - o["logLik"]) <= REL * abs(o["logLik"])
        assert np.abs(g["prob"] - o["prob"]).max() <= 1e-9
    assert g["logLik"] == pytest.approx(-1854.0001, abs=2e-4)      # survey known answer


def test_em_c2(gpu_handle, c2_donor):
    cb = _cb()
    A, D, Z, _ = c2_donor
    psi = np.array([0.4, 0.3, 0.2, 0.1])
    g = cb.clone_id_EM(A, D, Z, Psi=psi, theta_init=(0.05, 0.4), handle=gpu_handle, verbose=False)
    o = O.clone_id_EM(A, D, Z, Psi=psi, theta_init=(0.05, 0.4))
    assert g["n_iter"] == o["n_iter"]
    assert np.allclose(g["theta"], o["theta"], rtol=REL)
    assert abs(g["logLik"] - o["logLik"]) <= REL * abs(o["logLik"])
    assert np.abs(g["prob"] - o["prob"]).max() <= 1e-9


def test_get_loglik(gpu_handle, example_donor, c2_donor):
    cb = _cb()
    rng = np.random.default_rng(0)
    for A, D, Z in (example_donor, c2_donor[:3]):
        A1, B1, _ = O.preprocess(A, D)
        N, M = A.shape
        K = Z.shape[1]
        gpu_handle.load_dense(A, D)
        lab = rng.integers(1, K + 1, M)
        th1 = rng.uniform(0.2, 0.8, N)
        g = gpu_handle.get_loglik(Z, lab, 0.03, th1)
        o = O.get_logLik(A1, B1, Z, lab, 0.03, np.repeat(th1[:, None], M, 1))
        assert abs(g - o) <= REL * abs(o)
        P = rng.dirichlet(np.ones(K), M)
        C = rng.uniform(0, 1, Z.shape)
        g = gpu_handle.get_loglik(C, P, 0.03, th1)
        o = O.get_logLik(A1, B1, C, P, 0.03, np.repeat(th1[:, None], M, 1))
        assert abs(g - o) <= REL * abs(o)
        g = gpu_handle.get_loglik(C, P, 0.03, 0.41)                  # scalar theta1 (wise = global)
        o = O.get_logLik(A1, B1, C, P, 0.03, np.full((N, M), 0.41))
        assert abs(g - o) <= REL * abs(o)


@pytest.mark.parametrize("kw", [
    dict(),                                                   # the reference test's configuration
    dict(wise="global", relax_rate_fixed=0.3),
    dict(relax_Config=False),
    dict(keep_base_clone=False, relax_rate_prior=(2, 5)),
    dict(Psi=np.array([0.1, 0.2, 0.3, 0.4])),
])
def test_gibbs_stepwise_example_donor(gpu_handle, example_donor, kw):
    A, D, Z = example_donor
    res, v = _verify(_cb(), gpu_handle, A, D, Z, T=60, **kw)
    assert res["prob_all"].shape == (60, 428 * 4) and np.all(res["prob_all"][0] == 0)
    assert res["Config_all"].shape == (60, 34 * 4)
    assert res["prob"].shape == (428, 4) and res["prob_variant"].shape == (34, 428)


@pytest.mark.parametrize("K", [2, 3, 6, 12, 20])
def test_gibbs_denovo_clone_counts(gpu_handle, example_donor, K):
    """De-novo mode (Config = 0, relax rate fixed at 0.5, R/clone_id.R:98-105): exercises every padded
    clone width (4, 8, 16, 32) and the exact-tie path of the categorical draw."""
    A, D, _ = example_donor
    _verify(_cb(), gpu_handle, A, D, np.zeros((34, K)), T=30, relax_rate_fixed=0.5)


def test_gibbs_c2_four_chains(gpu_handle, c2_donor):
    """BASELINE config 2: sim_read_count donor, 4 clones, relax_Config = TRUE, n_chain = 4."""
    cb = _cb()
    A, D, Z, lab = c2_donor
    T = 40
    outs = cb.clone_id_Gibbs(A, D, Z, min_iter=T, max_iter=T, seed=5, n_chain=4, keep_assign_all=True,
                             handle=gpu_handle, verbose=False, _all_chains=True)
    assert len(outs) == 4
    for c, res in enumerate(outs):
        v = O.verify_gibbs_trace(A, D, Z, res, min_iter=T, seed=5, chain=c)
        assert v["max_dprob"] <= 1e-9 and v["hard"] == 0 and v["config_hard"] == 0, v
        assert v["max_dloglik_rel"] <= REL
    assert not np.array_equal(outs[0]["assign_all"], outs[1]["assign_all"])     # chains are independent
    merged = cb.clone_id(A, D, Z, min_iter=T, max_iter=T, seed=5, n_chain=4, handle=gpu_handle, verbose=False)
    assert merged["n_chain"] == 4 and np.allclose(merged["prob"].sum(1), 1)
    assert (merged["prob"].argmax(1) + 1 == lab).mean() > 0.8


@pytest.fixture()
def rows_handle(gpu_handle):
    """The session handle forced onto the three-kernel sweep with the row-split cell kernel (small donors take the
    single-CTA path of small.cu by default), restored afterwards."""
    gpu_handle.set_cells_layout("rows")
    yield gpu_handle
    gpu_handle.set_cells_layout("auto")


@pytest.mark.parametrize("kw", [dict(), dict(wise="global", relax_rate_fixed=0.3), dict(relax_Config=False),
                                dict(Psi=np.array([0.1, 0.2, 0.3, 0.4]), keep_base_clone=False)])
def test_three_kernel_path_stepwise(rows_handle, example_donor, kw):
    A, D, Z = example_donor
    _verify(_cb(), rows_handle, A, D, Z, T=40, **kw)


@pytest.mark.parametrize("K", [3, 12, 20])
def test_three_kernel_path_denovo(rows_handle, example_donor, K):
    A, D, _ = example_donor
    _verify(_cb(), rows_handle, A, D, np.zeros((34, K)), T=25, relax_rate_fixed=0.5)


def test_small_path_equals_three_kernel_path(gpu_handle, c2_donor):
    """Same Philox keys, same Beta sampler: the single-CTA sweep and the three-kernel sweep produce the same chain
    (labels, Config flips, theta draws; probabilities to rounding), four chains at once."""
    cb = _cb()
    A, D, Z, _ = c2_donor
    out = {}
    for layout in ("small", "rows"):
        gpu_handle.set_cells_layout(layout)
        out[layout] = cb.clone_id_Gibbs(A, D, Z, min_iter=100, max_iter=300, seed=5, n_chain=4, keep_assign_all=True,
                                        handle=gpu_handle, verbose=False, _all_chains=True)
    gpu_handle.set_cells_layout("auto")
    for a, b in zip(out["small"], out["rows"]):
        assert a["n_iter"] == b["n_iter"]
        assert np.array_equal(a["assign_all"], b["assign_all"]) and np.array_equal(a["Config_all"], b["Config_all"])
        assert np.abs(a["prob_all"] - b["prob_all"]).max() <= 1e-12
        assert np.allclose(a["theta1_all"], b["theta1_all"], rtol=1e-13, atol=0)
        assert np.allclose(a["logLik"], b["logLik"], rtol=1e-12, atol=0)
        assert np.abs(a["prob"] - b["prob"]).max() <= 1e-12 and a["percent_converged"] == b["percent_converged"]


@pytest.fixture()
def slices_handle(gpu_handle):
    """The session handle forced onto the slice-major cell kernel (one lane per cell), restored afterwards."""
    gpu_handle.set_cells_layout("slices")
    yield gpu_handle
    gpu_handle.set_cells_layout("auto")


@pytest.mark.parametrize("K", [0, 2, 3, 6, 12, 20])
def test_slice_kernel_stepwise(slices_handle, example_donor, K):
    """k_cells_sell against the oracle sweep by sweep: the reference test's configuration (K = 0 -> tree$Z)
    and de-novo runs over every padded clone width (packed masks for K <= 16, separate mask table for
    K > 16), including the exact-tie path (19 uncovered cells tie on all clones) and ragged slices
    (428 cells = 13 full slices + one of 12 lanes, rows of 0-12 entries)."""
    A, D, Z = example_donor
    if K == 0:
        _verify(_cb(), slices_handle, A, D, Z, T=60)
        _verify(_cb(), slices_handle, A, D, Z, T=30, wise="global", relax_rate_fixed=0.3,
                Psi=np.array([0.1, 0.2, 0.3, 0.4]))
    else:
        _verify(_cb(), slices_handle, A, D, np.zeros((34, K)), T=30, relax_rate_fixed=0.5)


def test_slice_kernel_c2_and_wide_index(slices_handle, c2_donor):
    A, D, Z, lab = c2_donor
    _verify(_cb(), slices_handle, A, D, Z, T=30, seed=5)
    test_many_variants_uint32_index_and_global_table(slices_handle)


def test_slice_kernel_equals_row_kernel_at_full_size(gpu_handle):
    """C3 shape: both cell kernels see the same state (teacher-forced by running one sweep from the same
    seed) and must agree on every probability to rounding and on (nearly) every label."""
    import bench
    N, M, K, cov = bench.WORKLOADS["C3"]
    cfg = bench.make_config(N, K)
    h = gpu_handle
    h.synth_generate(N, M, K, cfg, cov, 78)
    out = {}
    for layout in ("rows", "slices"):
        h.set_cells_layout(layout)
        h.gibbs_run(cfg, None, min_iter=3, max_iter=3, seed=6, keep_prob_all=1, keep_assign_all=1)
        out[layout] = (h.fetch("prob_all"), h.fetch("assign_all"), h.fetch("logLik"))
    h.set_cells_layout("auto")
    pr, zr, lr = out["rows"]
    ps, zs, ls = out["slices"]
    assert np.abs(pr[1] - ps[1]).max() <= 1e-9           # sweep 2 starts from identical state
    assert (zr[1] != zs[1]).mean() <= 1e-5
    if np.array_equal(zr[1], zs[1]):
        assert np.abs(pr[2] - ps[2]).max() <= 1e-9 and abs(lr[2] - ls[2]) <= 1e-9 * abs(lr[2])


@pytest.mark.parametrize("layout", ["rows", "slices"])
def test_mito_denovo_deep_counts(gpu_handle, layout):
    """The vignette's mtDNA case (vignettes/vignette-cloneid.Rmd:160-186): inst/extdata/passed_{ad,dp}.mtx,
    25 variants x 77 cells, 96 % dense, depths up to 6213, de-novo with 3 clones.  Per-variant depth totals
    exceed 65535, so k_stats must take its 64-bit bins; logits reach thousands."""
    import os
    from conftest import GOLDEN
    d = np.load(os.path.join(GOLDEN, "mito_passed.npz"))
    A, D = d["AD"].astype(float), d["DP"].astype(float)
    assert D.max() > 6000 and D.sum(1).max() > 65535
    gpu_handle.set_cells_layout(layout)
    cfg, T = np.zeros((25, 3)), 40
    try:
        res = _cb().clone_id_Gibbs(A, D, cfg, min_iter=T, max_iter=T, seed=13, keep_assign_all=True,
                                   handle=gpu_handle, verbose=False, relax_rate_fixed=0.5)
    finally:
        gpu_handle.set_cells_layout("auto")
    # the reference's dense get_logLik underflows to log(0) = -Inf at these depths;
    # the trace is checked against the table identity instead
    v = O.verify_gibbs_trace(A, D, cfg, res, min_iter=T, seed=13, relax_rate_fixed=0.5, dense_loglik=False)
    assert v["max_dprob"] <= 1e-9 and v["hard"] == 0 and v["config_hard"] == 0, v
    assert v["max_dloglik_rel"] <= REL, v


def test_both_statistic_kernels_agree(gpu_handle, monkeypatch):
    """k_stats_units (large shards) and k_stats_stream (small shards) compute the same integer tables: chains run
    under either are bit-identical.  70 000 cells = two cell blocks (one partial), ragged segments."""
    import bench
    N, M, K = 300, 70000, 6
    cfg = bench.make_config(N, K, seed=5)
    h = gpu_handle
    h.synth_generate(N, M, K, cfg, 0.08, 123)
    out = {}
    for kern in ("units", "stream"):
        monkeypatch.setenv("CDL_STATS_KERNEL", kern)
        h.gibbs_run(cfg, None, min_iter=8, max_iter=8, seed=2, keep_prob_all=0, keep_assign_all=1)
        out[kern] = (h.fetch("Config_all"), h.fetch("theta_traces"), h.fetch("assign_all"), h.fetch("logLik"))
    monkeypatch.delenv("CDL_STATS_KERNEL")
    a, b = out["units"], out["stream"]
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
    assert np.array_equal(a[1][0], b[1][0]) and np.array_equal(a[1][1], b[1][1])
    assert a[0][1:].any() and len(set(a[2][-1])) == K


@pytest.mark.parametrize("kw", [dict(), dict(relax_rate_fixed=0.3, keep_base_clone=False), dict(relax_Config=False)])
def test_element_wise_matches_oracle(gpu_handle, example_donor, kw):
    """wise = "element" (R/clone_id.R:413-416): one theta1 per covered entry.  The oracle re-runs the chain with
    the GPU's theta / relax-rate draws injected and its own uniforms from the same Philox streams: every
    probability, label, Config flip and the logLik trace must agree."""
    cb = _cb()
    A, D, Z = example_donor
    T, nnz = 40, 1874
    g = cb.clone_id_Gibbs(A, D, Z, wise="element", min_iter=T, max_iter=T, seed=21, keep_assign_all=True,
                          handle=gpu_handle, verbose=False, **kw)
    assert g["theta1_all"].shape == (T, nnz) and np.all((g["theta1_all"] > 0) & (g["theta1_all"] < 1))
    assert g["element"].shape == (nnz,) and np.array_equal(g["element"], np.flatnonzero(np.nan_to_num(D).T.ravel() > 0))
    assert np.isnan(g["theta1"]).sum() == 34 * 428 - nnz
    dr = O.ReplayDraws(seed=21, chain=0, theta0_all=g["theta0_all"], theta1_all=g["theta1_all"],
                       relax_rate_all=g["relax_rate_all"])
    o = O.clone_id_Gibbs(A, D, Z, wise="element", min_iter=T, max_iter=T, draws=dr, **kw)
    assert np.abs(g["prob_all"] - o["prob_all"]).max() <= 1e-9
    assert np.array_equal(g["assign_all"], o["assign_all"])
    assert np.array_equal(g["Config_all"], o["Config_all"])
    assert np.allclose(g["logLik"][1:], o["logLik"][1:], rtol=1e-9, atol=0)
    assert np.abs(g["prob"] - o["prob"]).max() <= 1e-9
    assert np.array_equal(np.isnan(g["prob_variant"]), np.isnan(o["prob_variant"]))
    assert np.nanmax(np.abs(g["prob_variant"] - o["prob_variant"])) <= 1e-9
    assert abs(g["logLik_post"] - o["logLik_post"]) <= REL * abs(o["logLik_post"])
    # the per-entry Beta draws themselves: posterior shape is prior + (A, B) where the entry's variant is in
    # the cell's clone, so covered entries of carried variants have larger means than the prior's 0.45
    assert 0.2 < g["theta1_all"][0].mean() < 0.7


def test_incremental_statistics_are_exact(gpu_handle, example_donor):
    """incremental_stats corrects the previous sweep's SA/SB table from the cells whose label changed (integer
    sums): chains are bit-identical to the full pass.  example_donor: most cells stay uncertain, so the ~3 %
    threshold sends nearly every sweep down the full pass; an informative synthetic donor: after a few sweeps only
    a handful of labels move and the incremental path carries the chain."""
    cb = _cb()
    A, D, Z = example_donor
    kw = dict(min_iter=40, max_iter=40, seed=8, keep_assign_all=True, handle=gpu_handle, verbose=False)
    gpu_handle.set_cells_layout("rows")          # both runs on the three-kernel sweep (the incremental update lives there)
    a = cb.clone_id_Gibbs(A, D, Z, incremental_stats=False, **kw)
    b = cb.clone_id_Gibbs(A, D, Z, incremental_stats=True, **kw)
    gpu_handle.set_cells_layout("auto")
    for f in ("prob_all", "assign_all", "Config_all", "theta1_all", "logLik"):
        assert np.array_equal(a[f], b[f]), f
    import bench
    for layout, (N, M, K, cov) in (("rows", (2000, 9000, 6, 0.15)), ("slices", (2000, 40000, 6, 0.15))):
        cfg = bench.make_config(N, K, seed=7)
        h = gpu_handle
        h.set_cells_layout(layout)
        h.synth_generate(N, M, K, cfg, cov, 55)
        out = {}
        T = 30
        for inc in (0, 1):
            h.gibbs_run(cfg, None, min_iter=T, max_iter=T, seed=3, keep_prob_all=0, keep_assign_all=1,
                        incremental_stats=inc)
            out[inc] = (h.fetch("assign_all"), h.fetch("Config_all"), h.fetch("theta_traces"), h.fetch("logLik"))
        h.set_cells_layout("auto")
        full = h.incremental_full_passes()
        assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
        assert np.array_equal(out[0][2][1], out[1][2][1]) and np.array_equal(out[0][3], out[1][3])
        changed = (out[0][0][2:] != out[0][0][1:-1]).sum(1)           # labels that moved, sweeps 3..T
        expect_full = 1 + int((changed > M // 32).sum())             # sweep 2 is forced, then by the threshold
        assert full == expect_full, (full, expect_full, changed)
        assert full < T - 1, changed                                  # the incremental path really ran


def test_replay_mode(gpu_handle, example_donor):
    """Correctness check 2 of the north star: inject a recorded run's uniforms and theta draws; the
    assignments and Config flips must be reproduced."""
    cb = _cb()
    A, D, Z = example_donor
    T, M, NK = 50, 428, 34 * 4
    rec = O.RecordingDraws(21, T, M, NK)
    o = O.clone_id_Gibbs(A, D, Z, min_iter=T, max_iter=T, draws=rec)
    replay = {"n_rows": T, "u_assign": rec.ua, "u_config": rec.uc, "theta0": o["theta0_all"].ravel(),
              "theta1": o["theta1_all"], "relax_rate": o["relax_rate_all"]}
    g = cb.clone_id_Gibbs(A, D, Z, min_iter=T, max_iter=T, replay=replay, keep_assign_all=True,
                          handle=gpu_handle, verbose=False)
    v = O.verify_gibbs_trace(A, D, Z, g, min_iter=T, u_assign=rec.ua, u_config=rec.uc)
    assert v["hard"] == 0 and v["config_hard"] == 0 and v["max_dprob"] <= 1e-9, v
    assert np.allclose(g["theta0_all"].ravel(), o["theta0_all"].ravel())
    # the free-running oracle chain itself is reproduced unless a rounding-order tie intervened
    if v["boundary"] + v["near_tie"] == 0:
        assert np.array_equal(g["assign_all"], o["assign_all"])
        assert np.array_equal(g["Config_all"], o["Config_all"])
        assert np.abs(g["prob"] - o["prob"]).max() <= 1e-9
        assert abs(g["DIC"]["DIC"] - o["DIC"]["DIC"]) <= 1e-6 * abs(o["DIC"]["DIC"])


def test_committed_golden_vectors(gpu_handle, example_donor):
    """The committed golden vectors (tests/golden/oracle_example_donor.npz, written by tools/make_golden.py from the
    oracle; a genuine R recording made with tools/record_reference.R has the same fields): the recorded uniforms
    and Beta draws injected through cdl_replay reproduce the recorded assignments and Config flips exactly and the
    probabilities / logLik trace to 1e-9; cdl_em_run and cdl_get_loglik reproduce the stored EM fixed point and
    log-likelihood to 1e-6 relative (the north star's bars)."""
    cb = _cb()
    g0 = np.load(os.path.join(ROOT, "tests", "golden", "oracle_example_donor.npz"))
    A, D, Z = example_donor
    T = int(g0["T"])
    replay = {"n_rows": T, "u_assign": g0["u_assign"], "u_config": g0["u_config"], "theta0": g0["theta0_all"],
              "theta1": g0["theta1_all"], "relax_rate": g0["relax_rate_all"]}
    for layout in ("auto", "rows"):
        gpu_handle.set_cells_layout(layout)
        try:
            g = cb.clone_id_Gibbs(A, D, Z, min_iter=T, max_iter=T, replay=replay, keep_assign_all=True,
                                  handle=gpu_handle, verbose=False)
        finally:
            gpu_handle.set_cells_layout("auto")
        assert np.array_equal(g["assign_all"], g0["assign_all"]) and np.array_equal(g["Config_all"], g0["Config_all"])
        assert np.abs(g["prob"] - g0["prob"]).max() <= 1e-9
        assert np.abs(g["prob_all"][T - 1] - g0["prob_all_last"]).max() <= 1e-9
        assert np.allclose(g["logLik"][1:], g0["logLik"][1:], rtol=REL, atol=0)
        assert np.array_equal(g["Config_prob"], g0["Config_prob"])
    em = cb.clone_id_EM(A, D, Z, theta_init=tuple(g0["em_theta_init"]), handle=gpu_handle, verbose=False)
    assert em["n_iter"] == int(g0["em_n_iter"]) and np.allclose(em["theta"], g0["em_theta"], rtol=REL, atol=0)
    assert abs(em["logLik"] - float(g0["em_logLik"])) <= REL * abs(float(g0["em_logLik"]))
    assert np.abs(em["prob"] - g0["em_prob"]).max() <= 1e-9
    it, K, N = int(g0["ll_it"]), Z.shape[1], Z.shape[0]
    ll = cb.get_logLik(A, D, g0["Config_all"][it - 1].reshape(K, N).T.astype(float), g0["assign_all"][it - 1],
                       float(g0["theta0_all"][it - 1]), g0["theta1_all"][it - 1], handle=gpu_handle)
    assert abs(ll - float(g0["ll_value"])) <= REL * abs(float(g0["ll_value"]))


def test_stop_rule_and_summaries(gpu_handle, example_donor):
    """Geweke stop rule (every 100 sweeps from min_iter), burn-in, posterior means, prob_variant, DIC."""
    cb = _cb()
    A, D, Z = example_donor
    g = cb.clone_id_Gibbs(A, D, Z, min_iter=100, max_iter=600, seed=3, handle=gpu_handle, verbose=False)
    it = g["n_iter"]
    assert it % 100 == 0 or it == 600
    for chk in range(100, it + 1, 100):
        frac = O.converged_fraction(g["prob_all"][:chk])
        if chk < it:
            assert not frac > 0.995          # the chain did not stop earlier
        elif it < 600:
            assert frac > 0.995              # and it stopped because the rule fired
    assert g["percent_converged"] == pytest.approx(O.r_round(O.converged_fraction(g["prob_all"][:it]), 3) * 100)
    nb = int(np.ceil(it * 0.5))
    assert g["n_buin"] == nb
    assert np.abs(g["prob"] - g["prob_all"][nb - 1:it].mean(0).reshape(4, 428).T).max() <= 1e-12
    assert np.abs(g["Config_prob"] - g["Config_all"][nb - 1:it].mean(0).reshape(4, 34).T).max() <= 1e-12
    assert g["theta0"] == pytest.approx(g["theta0_all"][nb - 1:it].mean())
    assert g["relax_rate"] == pytest.approx(g["relax_rate_all"][nb - 1:it].mean())
    # prob_variant (R/clone_id.R:606-622)
    from scipy.stats import binom
    A1, B1, _ = O.preprocess(A, D)
    p1 = sum(binom.pmf(A1, A1 + B1, g["theta1_all"][i][:, None]) for i in range(nb - 1, it))
    p0 = sum(binom.pmf(A1, A1 + B1, g["theta0_all"][i, 0]) for i in range(nb - 1, it))
    assert np.abs(g["prob_variant"] - p1 / (p1 + p0)).max() <= 1e-9
    th1m = np.repeat(g["theta1_all"][nb - 1:it].mean(0)[:, None], 428, 1)
    llp = O.get_logLik(A1, B1, g["Config_prob"], g["prob"], g["theta0"], th1m)
    assert abs(g["logLik_post"] - llp) <= REL * abs(llp)
    d = O.devianceIC(g["logLik"][nb - 1:it], llp)
    assert abs(g["DIC"]["DIC"] - d["DIC"]) <= 1e-6 * abs(d["DIC"])


def test_relabel(gpu_handle, example_donor):
    cb = _cb()
    A, D, Z = example_donor
    g = cb.clone_id_Gibbs(A, D, Z, min_iter=60, max_iter=60, seed=9, relabel=True, handle=gpu_handle,
                          verbose=False)
    assert g["prob"].shape == (428, 4) and np.allclose(g["prob"].sum(1), 1)


def test_reference_testthat_clone_id(gpu_handle, example_donor):
    """The reference's own hot-path test file, call for call (tests/testthat/test-clone_id.R:7-49): EM, the one
    sampler configuration it exercises (min_iter = 200, max_iter = 500, relax_Config, relabel), and the consumers
    of the result (assign_scores / multiPRC / binaryPRC, binaryROC).  Like there, the assertions are on types and
    shapes -- the reference pins nothing else."""
    cb = _cb()
    A_clone, D_clone, Z = example_donor
    assignments_EM = cb.clone_id(A_clone, D_clone, Config=Z, inference="EM", handle=gpu_handle, verbose=False)
    assert isinstance(assignments_EM, dict) and assignments_EM["prob"].shape == (428, 4)
    assignments = cb.clone_id(A_clone, D_clone, Config=Z, min_iter=200, max_iter=500, relax_Config=True,
                              relabel=True, handle=gpu_handle, verbose=False)
    assert isinstance(assignments, dict)
    assert assignments["prob"].shape == (428, 4) and assignments["prob_variant"].shape == (34, 428)
    assert assignments["Config_prob"].shape == (34, 4) and 200 <= assignments["n_iter"] <= 500
    I_sim = assignments["prob"] == cb.rowMax(assignments["prob"])[:, None]
    res = cb.assign_scores(assignments["prob"], I_sim, verbose=False)
    assert isinstance(res, dict) and set(res) == {"df_sg", "df_db", "AUC_sg", "AUC_db"}
    assert len(res["df_sg"]["cutoff"]) == 1001 and len(res["df_db"]["Recall"]) == 1001
    res = cb.binaryROC(assignments["prob"][:, 0], I_sim[:, 0])
    assert isinstance(res, dict) and 0.0 <= res["AUC"] <= 1.0


def test_converged_posterior_matches_oracle_chains(gpu_handle, c2_donor, example_donor):
    """Correctness check 3: converged prob agrees with independent CPU chains within Monte-Carlo error
    (max |dprob| <= 0.02, identical argmax where prob > 0.9) wherever the posterior is identified: the
    sim_read_count donor with relax_Config = TRUE (on the cells where chains mix, see below), and
    example_donor with the tree held fixed.
    (example_donor WITH relaxation is multimodal -- clone columns can swap -- so independent chains of the
    reference itself disagree there; it is covered step-wise above.)"""
    cb = _cb()
    A, D, Z, lab = c2_donor
    gs = cb.clone_id_Gibbs(A, D, Z, min_iter=1000, max_iter=1000, n_chain=4, seed=17, handle=gpu_handle,
                           verbose=False, keep_prob_all=0, _all_chains=True)
    os_ = [O.clone_id_Gibbs(A, D, Z, min_iter=500, max_iter=500, draws=O.NumpyDraws(100 + c)) for c in range(3)]
    pg = np.array([g["prob"] for g in gs])
    po_ = np.array([o["prob"] for o in os_])
    # With Config relaxation a handful of the 500 cells are bimodal (a variant flips and takes the cell with
    # it): independent chains OF THE ORACLE ITSELF then differ by 0.3-0.6 on those cells.  The comparison is
    # therefore |mean_gpu - mean_oracle| <= 0.02 + 4 standard errors of that difference, the standard error
    # estimated from the spread between chains -- i.e. the plain 0.02 bar wherever chains mix (which must be
    # the large majority of cells), and a z-test on the few cells where they do not.
    g_prob, po = pg.mean(0), po_.mean(0)
    se = np.sqrt(pg.var(0, ddof=1) / pg.shape[0] + po_.var(0, ddof=1) / po_.shape[0])
    assert np.all(np.abs(g_prob - po) <= 0.02 + 4 * se)
    tight = (4 * se).max(1) <= 0.02
    assert tight.mean() >= 0.8, tight.mean()
    sure = tight & (po.max(1) > 0.9)
    assert sure.sum() > 300 and np.array_equal(g_prob[sure].argmax(1), po[sure].argmax(1))
    # chain-level scalars: medians over chains (a chain that sits in the second mode has theta0 ~ 6x larger,
    # in the oracle as well)
    assert abs(np.median([g["theta0"] for g in gs]) - np.median([o["theta0"] for o in os_])) < 5e-4
    assert np.abs(np.median([g["Config_prob"] for g in gs], axis=0) -
                  np.median([o["Config_prob"] for o in os_], axis=0)).mean() < 0.02
    assert abs(np.median([g["relax_rate"] for g in gs]) - np.median([o["relax_rate"] for o in os_])) < 0.02
    # fixed tree on the real donor: the plain 0.02 bar.  Chain lengths are chosen so that Monte-Carlo error fits under it:
    # two groups of 2 x 2000-sweep oracle chains differ by up to 0.030 from each other, these four 4000-sweep oracle
    # chains are within 0.006 of the mean of sixteen more (measured on the host), 8 x 8000 sweeps here add ~0.004
    A, D, Z = example_donor
    g = cb.clone_id(A, D, Z, min_iter=8000, max_iter=8000, n_chain=8, seed=3, relax_Config=False,
                    handle=gpu_handle, verbose=False, keep_prob_all=0)
    os_ = [O.clone_id_Gibbs(A, D, Z, min_iter=4000, max_iter=4000, relax_Config=False,
                            draws=O.NumpyDraws(200 + c)) for c in range(4)]
    po = sum(o["prob"] for o in os_) / len(os_)
    assert np.abs(g["prob"] - po).max() <= 0.02
    sure = po.max(1) > 0.9
    assert np.array_equal(g["prob"][sure].argmax(1), po[sure].argmax(1))
    assert abs(g["theta0"] - np.mean([o["theta0"] for o in os_])) < 2e-3


@pytest.mark.parametrize("layout,buin_frac,snap32", [("rows", 0.5, 0), ("slices", 0.3, 1), ("rows", 0.5, 1)])
def test_streaming_stop_rule_matches_exact(gpu_handle, c2_donor, layout, buin_frac, snap32):
    """Streaming mode (no prob_all) applies the reference's Geweke stop rule and its burn-in, which depends on
    the final iteration (R/clone_id.R:580-606,645), from running column sums and window snapshots: same stop
    iteration, burn-in and posterior means as the exact mode that keeps the whole trace.  Snapshots kept as fp64
    reproduce the means to rounding; the default fp32 snapshots (half the memory) to ~1e-7."""
    cb = _cb()
    A, D, Z, _ = c2_donor
    gpu_handle.set_cells_layout(layout)
    try:
        out = {}
        for keep in (1, 0):
            out[keep] = cb.clone_id_Gibbs(A, D, Z, min_iter=300, max_iter=2500, seed=11, buin_frac=buin_frac,
                                          handle=gpu_handle, verbose=False, keep_prob_all=keep, snapshot_fp32=snap32)
    finally:
        gpu_handle.set_cells_layout("auto")
    ex, st = out[1], out[0]
    assert ex["exact_trace"] and not st["exact_trace"]
    assert st["n_iter"] == ex["n_iter"] and st["n_buin"] == ex["n_buin"]
    assert ex["n_iter"] < 2500, "pick parameters with which the exact run stops early"
    assert abs(st["percent_converged"] - ex["percent_converged"]) <= 0.2
    assert np.abs(st["prob"] - ex["prob"]).max() < (1e-6 if snap32 else 1e-9)   # difference of running sums vs mean of the rows
    assert np.array_equal(st["Config_prob"], ex["Config_prob"])
    assert st["theta0"] == ex["theta0"] and st["relax_rate"] == ex["relax_rate"]
    assert np.isclose(st["logLik_post"], ex["logLik_post"], rtol=1e-6 if snap32 else 1e-9)


def test_device_beta_sampler_distribution(gpu_handle):
    """The on-device Beta draws (Marsaglia-Tsang gammas on Philox streams) replace R's rbeta (Cheng BB/BC):
    same distribution, different stream.  Kolmogorov-Smirnov against scipy for small, moderate and large
    shapes, using the initial theta1 draws (one per variant, R/clone_id.R:462)."""
    from scipy import stats
    cb = _cb()
    N, M = 40000, 4
    rng = np.random.default_rng(0)
    p = np.arange(M + 1, dtype=np.int64) * 3
    rows = np.tile(np.array([0, 1, 2], np.int32), M)
    counts = cb.SparseCounts(N, M, p, rows, np.ones(rows.size, np.uint16), np.full(rows.size, 2, np.uint16))
    cfg = np.zeros((N, 2))
    for a, b in [(0.45, 0.55), (0.2, 99.8), (5.3, 2.1), (3000.0, 9000.0)]:
        res = cb.clone_id_Gibbs(counts, None, cfg, min_iter=3, max_iter=3, seed=int(a * 100), prior1=(a, b),
                                handle=gpu_handle, verbose=False, light=True)
        draws = res["theta1_all"][0]
        assert draws.shape == (N,)
        ks = stats.kstest(draws, stats.beta(a, b).cdf)
        assert ks.pvalue > 1e-3, (a, b, ks)
        assert abs(draws.mean() - a / (a + b)) < 5 * np.sqrt(a * b / ((a + b) ** 2 * (a + b + 1)) / N) + 1e-12
        # sweep 2 draws are posterior draws: for untouched variants (3..N-1) still Beta(a, b)
        ks2 = stats.kstest(res["theta1_all"][1][3:], stats.beta(a, b).cdf)
        assert ks2.pvalue > 1e-3, (a, b, ks2)


def test_sparse_input_equals_dense(gpu_handle, c2_donor):
    import scipy.sparse as sp
    cb = _cb()
    A, D, Z, _ = c2_donor
    Dn, An = np.nan_to_num(D), np.nan_to_num(A)
    An[Dn == 0] = 0
    kw = dict(min_iter=15, max_iter=15, seed=2, handle=gpu_handle, verbose=False)
    a = cb.clone_id_Gibbs(A, D, Z, **kw)
    b = cb.clone_id_Gibbs(sp.csc_matrix(An), sp.csc_matrix(Dn), Z, **kw)
    assert np.array_equal(a["prob_all"], b["prob_all"]) and np.array_equal(a["Config_all"], b["Config_all"])
    gpu_handle.load_dense(A, D)
    p, i, av, dv = gpu_handle.export_csc_u16()
    c = cb.clone_id_Gibbs(cb.SparseCounts(A.shape[0], A.shape[1], p, i, av, dv), None, Z, **kw)
    assert np.array_equal(a["prob_all"], c["prob_all"]) and np.array_equal(a["logLik"], c["logLik"])


def test_many_variants_uint32_index_and_global_table(gpu_handle):
    """N > 65536: uint32 variant ids in the cell-major layout and the weight table read from global
    memory instead of shared memory."""
    cb = _cb()
    rng = np.random.default_rng(8)
    N, M, K = 70000, 60, 5
    cfg = (rng.random((N, K)) < 0.3).astype(float)
    cfg[:, 0] = 0
    nnz_per = 300
    rows = np.concatenate([np.sort(rng.choice(N, nnz_per, replace=False)) for _ in range(M)]).astype(np.int32)
    p = np.arange(M + 1, dtype=np.int64) * nnz_per
    d = rng.integers(1, 12, rows.size).astype(np.uint16)
    a = rng.binomial(d, 0.3).astype(np.uint16)
    counts = cb.SparseCounts(N, M, p, rows, a, d)
    T = 12
    res = cb.clone_id_Gibbs(counts, None, cfg, min_iter=T, max_iter=T, seed=4, keep_assign_all=True,
                            handle=gpu_handle, verbose=False, light=False, keep_prob_all=1)
    import scipy.sparse as sp
    Dm = sp.csc_matrix((d.astype(float), rows, p), shape=(N, M)).toarray()
    Am = sp.csc_matrix((a.astype(float), rows, p), shape=(N, M)).toarray()
    Dm[Dm == 0] = np.nan
    Am[np.isnan(Dm)] = np.nan
    v = O.verify_gibbs_trace(Am, Dm, cfg, res, min_iter=T, seed=4, dense_loglik=False)
    assert v["max_dprob"] <= 1e-9 and v["hard"] == 0 and v["config_hard"] == 0 and v["max_dloglik_rel"] <= REL, v


def test_full_size_properties(gpu_handle):
    """At BASELINE's single-GPU sizes (C3 shape) through size-independent properties: probabilities sum to
    one, the sufficient statistics conserve the read totals (via the logLik identity), the run is
    bit-reproducible, an export -> reload round trip reproduces it, and labels are recovered."""
    cb = _cb()
    import bench
    N, M, K, cov = bench.WORKLOADS["C3"]
    cfg = bench.make_config(N, K)
    h = gpu_handle
    h.synth_generate(N, M, K, cfg, cov, 77)
    _, _, nnz, W = h.data_info()
    assert abs(nnz / (N * M) - cov) < 0.01
    kw = dict(min_iter=12, max_iter=12, seed=6, keep_prob_all=1)
    h.gibbs_run(cfg, None, **kw)
    prob_all = h.fetch("prob_all")
    ll = h.fetch("logLik")
    assert np.allclose(prob_all[1:].reshape(11, K, M).sum(1), 1, atol=1e-12)
    assert np.all(np.isfinite(ll[1:])) and np.all(ll[1:] < W)
    t0, t1 = h.fetch("theta_traces")
    assert np.all((t0 > 0) & (t0 < 1)) and np.all((t1 > 0) & (t1 < 1))
    lab = h.synth_labels()
    prob = h.fetch("prob")
    assert (prob.argmax(1) + 1 == lab).mean() > 0.98
    # scored like a simulation in the reference (R/assessment.R:354-373): clone columns match the truth, and the
    # precision / recall curve of the assignment is near 1
    cb = _cb()
    I_sim = np.zeros((M, K))
    I_sim[np.arange(M), lab - 1] = 1
    assert np.array_equal(cb.colMatch(I_sim, prob), np.arange(K))
    assert cb.multiPRC(prob, I_sim, verbose=False)["AUC"] > 0.97
    h.gibbs_run(cfg, None, **kw)
    assert np.array_equal(prob_all, h.fetch("prob_all")) and np.array_equal(ll, h.fetch("logLik"))
    p, i, a, d = h.export_csc_u16()
    assert p[-1] == nnz and np.all(d > 0) and np.all(a <= d)
    h.load_csc_u16(N, M, p, i, a, d)
    h.gibbs_run(cfg, None, **kw)
    assert np.array_equal(prob_all, h.fetch("prob_all")) and np.array_equal(ll, h.fetch("logLik"))
    # get_logLik on labels through the general kernel equals the sweep's table-identity trace
    t0, t1 = h.fetch("theta_traces")
    h.gibbs_run(cfg, None, keep_assign_all=1, **kw)
    z = h.fetch("assign_all")
    C = h.fetch("Config_all")
    it = 7
    g = h.get_loglik(C[it - 1].reshape(K, N).T, z[it - 1], t0[it - 2], t1[it - 2])
    assert abs(g - ll[it - 1]) <= 1e-9 * abs(g)


@pytest.mark.parametrize("workload", ["C3", "C4s8"])
def test_full_size_sweep_against_c_oracle(gpu_handle, workload):
    """At BASELINE's shapes (C3: 2000 x 100 000 x 8; a 1/8 shard of C4: 10 000 x 125 000 x 16) one sweep of the CUDA
    path is recomputed by the sparse C oracle from the same state (theta of trace row 1, input Config) and the
    same Philox uniforms: every probability to 1e-9, every label (unless its uniform sits on a CDF boundary),
    every Config flip and the logLik of the sweep."""
    import bench
    from oracle import c_oracle as CO
    from oracle import philox
    N, M, K, cov = bench.WORKLOADS[workload]
    cfg = bench.make_config(N, K)
    h = gpu_handle
    h.synth_generate(N, M, K, cfg, cov, 31)
    _, _, nnz, W = h.data_info()
    p, vi, a, d = h.export_csc_u16()
    seed = 12
    h.gibbs_run(cfg, None, min_iter=3, max_iter=3, seed=seed, keep_prob_all=1, keep_assign_all=1)
    t0, t1 = h.fetch("theta_traces")
    prob_g = h.fetch("prob_all")[1].reshape(K, M).T
    z_g = h.fetch("assign_all")[1] - 1
    cfg_g = h.fetch("Config_all")[1].reshape(K, N).T
    ll_g = h.fetch("logLik")[1]
    u = philox.uniforms(seed, 0, 2, philox.STREAM_ASSIGN, np.arange(M))
    prob_c, z_c, SA, SB = CO.sweep_cells(N, M, K, p, vi, a, d, CO.config_masks(cfg), t0[0], t1[0], np.full(K, 1 / K), u)
    assert np.abs(prob_g - prob_c).max() <= 1e-9
    bad = np.flatnonzero(z_g != z_c)
    assert bad.size <= 2
    for j in bad:                                    # only a uniform within 1e-6 of a CDF boundary may differ
        cdf = np.cumsum(np.sort(prob_c[j])[::-1])
        assert np.abs(cdf - u[j]).min() <= 1e-6
    if bad.size == 0:
        uc = philox.uniforms(seed, 0, 2, philox.STREAM_CONFIG, np.arange(N * K))
        mask_new, S1, S2, S3, S4, ll_c, diff1 = CO.sweep_table(SA, SB, CO.config_masks(cfg), True, True, 0.1,
                                                               t0[0], t1[0], uc, W)
        assert np.array_equal(CO.masks_to_config(mask_new, K), cfg_g)
        assert abs(ll_g - ll_c) <= 1e-9 * abs(ll_c)


@pytest.mark.parametrize("layout", ["auto", "rows", "slices"])
def test_edge_shapes(gpu_handle, layout):
    """Ragged and degenerate inputs on every cell kernel: cells without any covered variant, a variant nobody
    covers, a single cell, a single variant, K = 2 and K = 32, depth-1 entries only."""
    cb = _cb()
    rng = np.random.default_rng(3)
    gpu_handle.set_cells_layout(layout)
    try:
        for N, M, K in ((1, 40, 2), (7, 1, 3), (9, 70, 32), (40, 33, 2)):
            D = rng.integers(0, 4, (N, M)).astype(float)
            D[:, ::5] = 0                                   # every fifth cell has no data at all
            if N > 2:
                D[2, :] = 0                                 # a variant without any coverage
            A = rng.binomial(D.astype(int), 0.3).astype(float)
            D[D == 0] = np.nan
            A[np.isnan(D)] = np.nan
            cfg = (rng.random((N, K)) < 0.4).astype(float)
            cfg[:, 0] = 0
            if np.isnan(D).all():
                continue
            T = 12
            res = cb.clone_id_Gibbs(A, D, cfg, min_iter=T, max_iter=T, seed=5, keep_assign_all=True,
                                    handle=gpu_handle, verbose=False)
            v = O.verify_gibbs_trace(A, D, cfg, res, min_iter=T, seed=5)
            assert v["max_dprob"] <= 1e-9 and v["hard"] == 0 and v["config_hard"] == 0, (N, M, K, v)
            assert v["max_dloglik_rel"] <= REL or abs(res["logLik"][-1]) < 1e-9, (N, M, K, v)
            assert np.allclose(res["prob"].sum(1), 1)
    finally:
        gpu_handle.set_cells_layout("auto")


def test_error_paths(gpu_handle, example_donor):
    cb = _cb()
    A, D, Z = example_donor
    with pytest.raises(cb.CdlError, match="non-integer"):
        B = A.copy(); B[np.isfinite(B)] += 0.5
        gpu_handle.load_dense(B, D)
    with pytest.raises(cb.CdlError, match="exceeds"):
        B = np.where(np.isfinite(D), D + 1, np.nan)
        gpu_handle.load_dense(B, D)
    with pytest.raises(cb.CdlError, match="65535"):
        gpu_handle.load_dense(np.zeros((2, 2)), np.full((2, 2), 70000.0))
    gpu_handle.load_dense(A, D)
    with pytest.raises(cb.CdlError, match="1..32"):
        gpu_handle.gibbs_run(np.zeros((34, 40)), None, min_iter=5, max_iter=5)
    with pytest.raises(cb.CdlError, match="binary"):
        gpu_handle.gibbs_run(np.full((34, 3), 0.5), None, min_iter=5, max_iter=5)
    with pytest.raises(ValueError, match="relax_rate_fixed"):
        cb.clone_id_Gibbs(A, D, Z, relax_rate_fixed=1.5, handle=gpu_handle)
    with pytest.raises(ValueError, match="wise"):
        cb.clone_id_Gibbs(A, D, Z, wise="cell", handle=gpu_handle)
    with pytest.raises(ValueError, match="same size"):
        cb.clone_id_Gibbs(A, D[:, :-1], Z, handle=gpu_handle)
