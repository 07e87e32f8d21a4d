"""GPU parity tests added in round 2 (run with -m gpu on a B200): cell shards on one GPU, several consecutive sweeps at
BASELINE's full sizes against the sparse C oracle including the integers that enter the Beta draws, relabel against the
oracle, the vignette call, the streaming stop rule at saturated probabilities and at 10^6 cells, deep-count
prob_variant, and the device generator's distributions."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from oracle import cardelino_oracle as O  # noqa: E402
from oracle import philox  # noqa: E402

REL = 1e-6


def _cb():
    import cardelino_b200 as cb
    return cb


# ---------------------------------------------------------------------------------------------------------------------
def test_cell_shards_on_one_gpu(gpu_handle):
    """The multi-GPU partition without a second GPU: the donor is cut into two cell shards (cdl_set_shard), each
    loaded into its own handle on cuda:0.  Philox streams are keyed by the GLOBAL cell id, so the first sweep's labels
    of the shards are the unsharded run's labels, and the shards' SA/SB statistic tables (test hook, taken before the
    table kernel) add up to the unsharded table -- the sum the ranks exchange per sweep (R/clone_id.R:537-562)."""
    cb = _cb()
    import bench
    N, M, K, cov = 1500, 40000, 6, 0.1
    cfg = bench.make_config(N, K, seed=3)
    kw = dict(min_iter=2, max_iter=2, seed=5, keep_prob_all=1, keep_assign_all=1, incremental_stats=0)
    h = gpu_handle
    h.synth_generate(N, M, K, cfg, cov, 91)
    h.debug_capture(True)
    try:
        h.gibbs_run(cfg, None, **kw)
        z_full = h.fetch("assign_all")[1]
        p_full = h.fetch("prob_all")[1].reshape(K, M).T
        SA, SB, _, _, _ = h.debug_fetch()
    finally:
        h.debug_capture(False)
    assert SA.sum() > 0 and len(set(z_full)) == K
    cut = 17003                                   # not a multiple of 32 / 65536: ragged shards
    sa, sb, z_parts, p_parts = np.zeros_like(SA), np.zeros_like(SB), [], []
    for off, ml in ((0, cut), (cut, M - cut)):
        h2 = cb.Handle(0)
        try:
            h2.synth_generate(N, M, K, cfg, cov, 91, cell_offset=off, M_local=ml)     # sets the shard bounds itself
            h2.debug_capture(True)
            h2.gibbs_run(cfg, None, **kw)
            z_parts.append(h2.fetch("assign_all")[1])
            p_parts.append(h2.fetch("prob_all")[1].reshape(K, ml).T)
            a, b, _, _, _ = h2.debug_fetch()
            sa += a
            sb += b
        finally:
            h2.close()
    assert np.array_equal(np.concatenate(z_parts), z_full)
    assert np.abs(np.concatenate(p_parts) - p_full).max() <= 1e-12       # other slices, other summation order
    assert np.array_equal(sa, SA) and np.array_equal(sb, SB)
    # the same through the loader: export a shard, load it into a fresh handle and declare it with cdl_set_shard
    p, vi, a, d = h.export_csc_u16()
    h3 = cb.Handle(0)
    try:
        lo, hi = p[cut], p[M]
        counts = cb.SparseCounts(N, M - cut, p[cut:] - lo, vi[lo:hi], a[lo:hi], d[lo:hi], cell_offset=cut, M_total=M)
        g = cb.clone_id_Gibbs(counts, None, cfg, min_iter=2, max_iter=2, seed=5, keep_assign_all=True, handle=h3,
                              verbose=False, keep_prob_all=1)
        assert np.array_equal(g["assign_all"][1], z_full[cut:])
    finally:
        h3.close()


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("workload", ["C3", "C4"])
def test_full_size_consecutive_sweeps_against_c_oracle(gpu_handle, workload):
    """BASELINE's full shapes (C3: 2000 x 100 000 x 8; C4: 10 000 x 1 000 000 x 16, the metric's configuration), the
    first SEVEN sweeps of a chain -- flipped Config from sweep 3 on, the learned relax rate in sweeps 7 and 8 -- each
    recomputed by the sparse C oracle from the previous row of the GPU's own traces and the same Philox uniforms:
    every probability to 1e-9, every label, the N x K read sums SA / SB exactly, every Config flip, the logLik of the
    sweep, and the integers that enter the Beta draws (S1..S4 and diff1, R/clone_id.R:502-504,546-571) EXACTLY; the
    drawn theta0 / theta1 / relax rate then follow from those shapes through the sampler's own Philox streams."""
    import bench
    from oracle import c_oracle as CO
    CO.set_threads()
    N, M, K, cov = bench.WORKLOADS[workload]
    cfg = bench.make_config(N, K)
    cfg_mask = CO.config_masks(cfg)
    h = gpu_handle
    h.synth_generate(N, M, K, cfg, cov, 31)
    _, _, nnz, W = h.data_info()
    p, vi, a, d = h.export_csc_u16()
    seed, T = 12, 8
    kw = dict(min_iter=2, seed=seed, keep_assign_all=1)         # min_iter = 2: the relax rate is learned from it = 6 on
    h.debug_capture(True)
    try:
        caps = {}
        for t in range(2, T + 1):                               # the hook keeps the last sweep: one run per length
            h.gibbs_run(cfg, None, max_iter=t, keep_prob_all=1 if t == T else 0, **kw)
            caps[t] = h.debug_fetch()
        t0, t1 = h.fetch("theta_traces")
        prob_all = h.fetch("prob_all")
        z_all = h.fetch("assign_all")
        cfg_all = h.fetch("Config_all")
        ll = h.fetch("logLik")
        _, relax_all = h.fetch("relax_rate")
    finally:
        h.debug_capture(False)
    Psi = np.full(K, 1.0 / K)
    mask_prev = cfg_mask
    learned = 0
    for it in range(2, T + 1):
        u = philox.uniforms(seed, 0, it, philox.STREAM_ASSIGN, np.arange(M))
        prob_c, z_c, SA, SB = CO.sweep_cells(N, M, K, p, vi, a, d, mask_prev, t0[it - 2], t1[it - 2], Psi, u)
        prob_g = prob_all[it - 1].reshape(K, M).T
        assert np.abs(prob_g - prob_c).max() <= 1e-9, it
        assert np.array_equal(z_all[it - 1] - 1, z_c), (it, int((z_all[it - 1] - 1 != z_c).sum()))
        gSA, gSB, gS3, gS4, gtot = caps[it]
        assert np.array_equal(gSA, SA) and np.array_equal(gSB, SB), it
        # relax rate used by this sweep: 0.1 (prior mean) until the reference starts learning it (:501-517)
        rate = relax_all[it - 1] if relax_all[it - 1] > 0 else 0.1
        learned += relax_all[it - 1] > 0
        uc = philox.uniforms(seed, 0, it, philox.STREAM_CONFIG, np.arange(N * K))
        mask_new, S1, S2, S3, S4, ll_c, diff1 = CO.sweep_table(SA, SB, cfg_mask, True, True, rate, t0[it - 2], t1[it - 2], uc, W)
        assert np.array_equal(CO.masks_to_config(mask_new, K), cfg_all[it - 1].reshape(K, N).T), it
        assert abs(ll[it - 1] - ll_c) <= 1e-9 * abs(ll_c), it
        assert np.array_equal(gS3.astype(np.float64), S3) and np.array_equal(gS4.astype(np.float64), S4), it
        assert [int(x) for x in gtot] == [int(S1), int(S2), int(S3.sum()), int(S4.sum()), int(diff1)], it
        # the draws from those shapes (:563-571), same Philox streams; libm differs in the last bits and a rejection
        # test may very rarely fall the other way
        th1 = philox.beta_parallel(0.45 + S3, 0.55 + S4, seed, 0, it, philox.STREAM_THETA1, np.arange(N))
        rel = np.abs(th1 - t1[it - 1]) / t1[it - 1]
        assert np.mean(rel <= 1e-9) >= 0.999 and np.median(rel) <= 1e-13, (it, float(np.mean(rel <= 1e-9)))
        th0 = philox.beta_parallel(np.array([0.2 + S1]), np.array([99.8 + S2]), seed, 0, it, philox.STREAM_THETA0, np.array([0]))
        assert abs(th0[0] - t0[it - 1]) <= 1e-9 * th0[0], it
        if it < T and relax_all[it] > 0:
            rr = philox.beta_parallel(np.array([1.0 + diff1]), np.array([9.0 + N * (K - 1) - diff1]), seed, 0, it + 1,
                                      philox.STREAM_RELAX, np.array([0]))
            assert abs(rr[0] - relax_all[it]) <= 1e-9 * rr[0], it
        mask_prev = mask_new
    assert learned >= 2 and not np.array_equal(mask_prev, cfg_mask)      # learned rate and flipped Config were exercised


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("K", [0, 6])
def test_relabel_matches_oracle(gpu_handle, example_donor, K):
    """relabel = TRUE (R/clone_id.R:624-644), done on the device: the oracle replays the same chain (the GPU's theta /
    relax draws injected, the same Philox uniforms) and relabels it with its own restatement of the loop -- K = 4 takes
    colMatch(force = TRUE) over all permutations, the de-novo K = 6 chain the greedy branch."""
    cb = _cb()
    A, D, Z = example_donor
    kw = {}
    if K:
        Z = np.zeros((A.shape[0], K))
        kw = dict(relax_rate_fixed=0.5, buin_frac=0.05)        # relabel from row 3 on: the early rows still swap clones
    T = 60
    g = cb.clone_id_Gibbs(A, D, Z, min_iter=T, max_iter=T, seed=9, relabel=True, handle=gpu_handle, verbose=False, **kw)
    dr = O.ReplayDraws(seed=9, chain=0, theta0_all=g["theta0_all"], theta1_all=g["theta1_all"],
                       relax_rate_all=g["relax_rate_all"])
    o = O.clone_id_Gibbs(A, D, Z, min_iter=T, max_iter=T, draws=dr, relabel=True, **kw)
    raw = cb.clone_id_Gibbs(A, D, Z, min_iter=T, max_iter=T, seed=9, relabel=False, handle=gpu_handle, verbose=False, **kw)
    assert np.abs(g["prob_all"] - o["prob_all"]).max() <= 1e-9
    assert np.array_equal(g["Config_all"], o["Config_all"])
    assert np.abs(g["prob"] - o["prob"]).max() <= 1e-9 and np.abs(g["Config_prob"] - o["Config_prob"]).max() <= 1e-12
    assert abs(g["logLik_post"] - o["logLik_post"]) <= REL * abs(o["logLik_post"])
    assert abs(g["DIC"]["DIC"] - o["DIC"]["DIC"]) <= 1e-6 * abs(o["DIC"]["DIC"])
    nb = g["n_buin"]
    assert np.array_equal(g["prob_all"][:nb - 1], raw["prob_all"][:nb - 1])       # rows before n_buin are untouched
    # whether the relabelling moved any column is a property of the chain: the same answer as the oracle's loop
    o_raw = O.clone_id_Gibbs(A, D, Z, min_iter=T, max_iter=T, draws=dr, relabel=False, **kw)
    assert np.array_equal(g["prob_all"], raw["prob_all"]) == np.array_equal(o["prob_all"], o_raw["prob_all"])


# ---------------------------------------------------------------------------------------------------------------------
def test_vignette_cellsnp_call(gpu_handle):
    """The vignette's configuration (vignettes/vignette-cloneid.Rmd:66-113): cellSNP VCF (112 SNVs x 77 cells, decoded
    by tools/vcf_to_npz.py) + the Canopy tree (3 clones) through clone_id(min_iter = 800, max_iter = 1200), variant
    ids matched by name like R/clone_id.R:104-118.  Step-wise against the oracle, then the converged posterior against
    independent oracle chains (north star check 3: max |dprob| <= 0.02 wherever the chains agree among themselves,
    same argmax where prob > 0.9)."""
    import pandas as pd
    cb = _cb()
    v = np.load(os.path.join(ROOT, "tests", "golden", "vignette_cellsnp.npz"))
    t = np.load(os.path.join(ROOT, "tests", "golden", "canopy_tree.npz"))
    A = pd.DataFrame(v["A"], index=list(v["variants"]), columns=list(v["cells"]))
    D = pd.DataFrame(v["D"], index=list(v["variants"]), columns=list(v["cells"]))
    ids = [s.replace(":", "_") for s in t["variants"]]                      # the vignette's gsub(":", "_", ...)
    perm = np.random.default_rng(0).permutation(len(ids))                  # Config rows in another order than D's
    Config = pd.DataFrame(t["Z"][perm], index=[ids[q] for q in perm], columns=list(t["clones"]))
    res = cb.clone_id(A, D, Config=Config, min_iter=800, max_iter=1200, seed=7, handle=gpu_handle, verbose=False,
                      keep_assign_all=True)
    assert res["prob"].shape == (77, 3) and res["prob_variant"].shape == (112, 77) and 800 <= res["n_iter"] <= 1200
    assert res["clone_names"] == ["clone1", "clone2", "clone3"] and len(res["cell_names"]) == 77
    # rows are in Config's order (intersect(rownames(Config), rownames(D))): the oracle gets the same
    order = [list(v["variants"]).index(s) for s in Config.index]
    Am, Dm, Z = v["A"][order], v["D"][order], Config.to_numpy()
    chk = O.verify_gibbs_trace(Am, Dm, Z, res, min_iter=800, seed=7)
    assert chk["max_dprob"] <= 1e-9 and chk["hard"] == 0 and chk["config_hard"] == 0 and chk["max_dloglik_rel"] <= REL, chk
    gs = cb.clone_id_Gibbs(Am, Dm, Z, min_iter=3000, max_iter=3000, n_chain=4, seed=70, handle=gpu_handle, verbose=False,
                           keep_prob_all=0, _all_chains=True)
    os_ = [O.clone_id_Gibbs(Am, Dm, Z, min_iter=1500, max_iter=1500, draws=O.NumpyDraws(300 + c)) for c in range(3)]
    pg, po = np.array([g["prob"] for g in gs]), np.array([o["prob"] for o in os_])
    se = np.sqrt(pg.var(0, ddof=1) / pg.shape[0] + po.var(0, ddof=1) / po.shape[0])
    assert np.all(np.abs(pg.mean(0) - po.mean(0)) <= 0.02 + 4 * se)
    sure = ((4 * se).max(1) <= 0.02) & (po.mean(0).max(1) > 0.9)
    assert np.array_equal(pg.mean(0)[sure].argmax(1), po.mean(0)[sure].argmax(1))
    df = cb.assign_cells_to_clones(res["prob"], cell_names=res["cell_names"], clone_names=res["clone_names"])
    assert len(df["clone"]) == 77 and set(df["clone"]) <= {"clone1", "clone2", "clone3", "unassigned"}


# ---------------------------------------------------------------------------------------------------------------------
def test_streaming_stop_rule_saturated_probabilities(gpu_handle):
    """The streaming stop rule (running sums of prob and prob^2, one-pass variances) on a donor whose cells are
    confidently assigned -- columns of prob_all that sit at 1 - 1e-10 or 1e-30 for thousands of rows, where a one-pass
    variance is pure cancellation.  It must still give the exact rule's verdicts, because window A contains trace
    row 1 (all zero in the reference, R/clone_id.R:426,466), which keeps vA >= ~m^2 / nA: same converged fraction at
    every check, same stop iteration as the two-pass kernel on the stored trace."""
    cb = _cb()
    import bench
    N, M, K, cov = 300, 6000, 5, 0.3                 # ~90 covered variants per cell: most columns saturate
    cfg = bench.make_config(N, K, seed=11)
    h = gpu_handle
    h.synth_generate(N, M, K, cfg, cov, 202)
    p, vi, a, d = h.export_csc_u16()
    counts = cb.SparseCounts(N, M, p, vi, a, d)
    saturated = None
    for min_iter in (20, 110, 700):                  # confident cells pass the test at the first check: one check per run,
        out, fr = {}, {}                             # so three runs with windows of 2, 11 and 70 rows
        for keep in (1, 0):
            fr[keep] = []
            out[keep] = cb.clone_id_Gibbs(counts, None, cfg, min_iter=min_iter, max_iter=min_iter + 300, seed=5, handle=h,
                                          verbose=False, keep_prob_all=keep, snapshot_fp32=0, light=True, check_every=10,
                                          progress=lambda c, it, f, k=keep: fr[k].append((it, f)) and False)
        ex, st = out[1], out[0]
        assert ex["exact_trace"] and not st["exact_trace"] and ex["stop_rule"] == 0 and st["stop_rule"] == 1
        checks_e = [(it, f) for it, f in fr[1] if f == f]
        checks_s = [(it, f) for it, f in fr[0] if f == f]
        assert len(checks_e) >= 1 and [c[0] for c in checks_e] == [c[0] for c in checks_s], (checks_e, checks_s)
        assert np.allclose([c[1] for c in checks_e], [c[1] for c in checks_s], atol=2e-4), (checks_e, checks_s)
        assert st["n_iter"] == ex["n_iter"] and st["n_buin"] == ex["n_buin"]
        assert abs(st["percent_converged"] - ex["percent_converged"]) <= 0.1
        assert np.abs(st["prob"] - ex["prob"]).max() < 1e-9
        saturated = np.mean((ex["prob"] > 1 - 1e-9) | (ex["prob"] < 1e-9))
    assert saturated > 0.5, saturated                # the regime the test is about


def test_streaming_stop_rule_at_a_million_cells(gpu_handle):
    """The reference's stop rule at >= 10^6 x K trace columns inside a 16 GB trace budget: the windowed Geweke test
    (fp32 snapshots of the running sums) stops the chain at the same iteration as the exact rule on the stored
    prob_all (which needs 10^6 x 4 x 700 doubles = 22 GB on the device), with the same burn-in and posterior."""
    cb = _cb()
    import bench
    N, M, K, cov = 200, 1000000, 4, 0.04             # ~8 covered variants per cell: cells stay uncertain, chains mix
    cfg = bench.make_config(N, K, seed=12)
    h = gpu_handle
    h.synth_generate(N, M, K, cfg, cov, 303)
    kw = dict(min_iter=300, max_iter=700, seed=6, check_every=100, relax_Config=1)
    h.gibbs_run(cfg, None, keep_prob_all=1, trace_budget_bytes=60 << 30, **kw)
    ie = h.gibbs_info()
    pe = h.fetch("prob")
    h.gibbs_run(cfg, None, keep_prob_all=0, trace_budget_bytes=16 << 30, **kw)
    is_ = h.gibbs_info()
    ps = h.fetch("prob")
    assert ie.exact_trace == 1 and is_.exact_trace == 0
    assert is_.n_iter == ie.n_iter and is_.n_buin == ie.n_buin, (is_.n_iter, ie.n_iter)
    assert abs(is_.percent_converged - ie.percent_converged) <= 0.1
    assert np.abs(ps - pe).max() < 1e-6
    assert is_.stop_rule == 1 and ie.stop_rule == 0
    # the reference's default schedule (5000 / 20000 / 100: 151 checks, ~250 snapshots alive at once) fits the same
    # budget on this donor, so a default clone_id() call keeps the reference's stop rule
    h.gibbs_run(cfg, None, keep_prob_all=0, trace_budget_bytes=16 << 30, seed=6)
    assert h.gibbs_info().stop_rule == 1 and 5000 <= h.gibbs_info().n_iter <= 20000


# ---------------------------------------------------------------------------------------------------------------------
def test_prob_variant_deep_counts_against_scipy(gpu_handle):
    """prob_variant (R/clone_id.R:606-622) on the vignette's mtDNA counts, where dbinom's two arms have exponents near
    -1300 (a = 326 of d = 6213): sum_t dbinom(a, d, th1_t) / (sum_t dbinom(.., th1_t) + sum_t dbinom(.., th0_t)) from
    scipy's log-pmf, max-shifted.  Every covered entry must be finite and agree; uncovered entries are 0.5."""
    from scipy import stats
    from scipy.special import logsumexp
    from conftest import GOLDEN
    d = np.load(os.path.join(GOLDEN, "mito_passed.npz"))
    A, D = d["AD"].astype(float), d["DP"].astype(float)
    T = 40
    g = _cb().clone_id_Gibbs(A, D, np.zeros((25, 3)), min_iter=T, max_iter=T, seed=13, handle=gpu_handle, verbose=False,
                             relax_rate_fixed=0.5)
    pv = g["prob_variant"]
    assert pv.shape == A.shape and np.all(np.isfinite(pv))
    nb = g["n_buin"]
    th0, th1 = g["theta0_all"][nb - 1:T, 0], g["theta1_all"][nb - 1:T]            # rows n_buin..it
    cov = D > 0
    l1 = logsumexp(stats.binom.logpmf(A[None], D[None], th1[:, :, None]), axis=0)
    l0 = logsumexp(stats.binom.logpmf(A[None], D[None], th0[:, None, None]), axis=0)
    with np.errstate(over="ignore"):
        want = np.where(cov, 1.0 / (1.0 + np.exp(l0 - l1)), 0.5)
    assert np.nanmax(np.abs(pv - want)) <= 1e-9
    deep = D > 3000
    assert deep.any() and np.all(np.isfinite(pv[deep]))


# ---------------------------------------------------------------------------------------------------------------------
def test_device_generator_follows_sim_read_count(gpu_handle):
    """The device generator (csrc/synth.cu) against the distributions of sim_read_count / sample_seq_depth
    (R/reads_simulator.R:53-152,174-212): labels ~ Multinomial(1/K) (:80), coverage ~ Bernoulli(cov) per entry, depths
    from the empirical table of the packaged D_input (:197-209), theta1 ~ Beta(0.45, 0.55) per variant and
    A ~ Binomial(D, theta1) where the cell's clone carries the variant (:95-109,125-129), theta0 ~ Beta(0.2, 99.8) per
    entry and A ~ Binomial(D, theta0) elsewhere (a Beta-binomial marginally)."""
    from scipy import stats
    from oracle import c_oracle as CO
    import bench
    N, M, K, cov = 1500, 60000, 6, 0.06
    cfg = bench.make_config(N, K, seed=21)
    h = gpu_handle
    h.synth_generate(N, M, K, cfg, cov, 404)
    p, vi, a, d = h.export_csc_u16()
    lab = h.synth_labels() - 1
    nnz = p[-1]
    # labels: chi-square against the uniform multinomial
    assert stats.chisquare(np.bincount(lab, minlength=K)).pvalue > 1e-4
    # coverage: per-cell counts ~ Binomial(N, cov)
    per_cell = np.diff(p)
    assert abs(per_cell.mean() - N * cov) < 4 * np.sqrt(N * cov * (1 - cov) / M)
    assert abs(per_cell.var() / (N * cov * (1 - cov)) - 1) < 0.05
    # depths: the 256-quantile table of D_input
    vals, cnt = np.unique(CO.DEPTH_Q, return_counts=True)
    obs = np.array([(d == v).sum() for v in vals])
    assert obs.sum() == nnz
    assert stats.chisquare(obs, cnt / 256.0 * nnz).pvalue > 1e-4
    # alternative reads where the clone carries the variant: per-variant rate = its theta1 ~ Beta(0.45, 0.55)
    cell = np.repeat(np.arange(M), per_cell)
    carried = cfg[vi, lab[cell]] > 0
    sa = np.bincount(vi[carried], weights=a[carried], minlength=N)
    sd = np.bincount(vi[carried], weights=d[carried], minlength=N)
    ok = sd > 2000                                                  # rates known to ~1 %
    assert ok.sum() > 500
    assert stats.kstest(sa[ok] / sd[ok], stats.beta(0.45, 0.55).cdf).pvalue > 1e-3
    # elsewhere: Beta-binomial(D, 0.2, 99.8); check P(A = 0 | D) for a few depths and the overall mean rate 0.002
    rest = ~carried
    assert abs(a[rest].sum() / d[rest].sum() - 0.002) < 2e-4
    for dep in (1, 5, 30):
        sel = rest & (d == dep)
        p0 = stats.betabinom(dep, 0.2, 99.8).pmf(0)
        n = sel.sum()
        assert abs((a[sel] == 0).mean() - p0) < 5 * np.sqrt(p0 * (1 - p0) / n) + 1e-4, dep


# ---------------------------------------------------------------------------------------------------------------------
def test_input_forms_and_argument_checks(gpu_handle, c2_donor):
    """Round-1 review items on the host interface: compact SparseCounts through the top-level clone_id() (with the chain
    merge), sparse A entries where D is zero or absent are dropped like A[which(D == 0)] <- NA (R/clone_id.R:380), a
    prior1 matrix must have one row per theta1 element (:418-424), relax_Config = 0 means FALSE."""
    import scipy.sparse as sp
    cb = _cb()
    A, D, Z, _ = c2_donor
    kw = dict(min_iter=30, max_iter=30, seed=4, handle=gpu_handle, verbose=False)
    dense = cb.clone_id(A, D, Config=Z, n_chain=2, **kw)
    counts = cb.SparseCounts.from_sparse(sp.csc_matrix(np.nan_to_num(A)), sp.csc_matrix(np.nan_to_num(D)))
    compact = cb.clone_id(counts, None, Config=Z, n_chain=2, **kw)
    assert compact["n_chain"] == 2 and np.array_equal(compact["prob"], dense["prob"])
    assert np.array_equal(compact["Config_prob"], dense["Config_prob"])
    # A has counts where D is zero / has no entry: dropped by every loader
    A2 = np.nan_to_num(A).copy()
    hole = np.argwhere(np.nan_to_num(D) == 0)[:50]
    A2[hole[:, 0], hole[:, 1]] = 3
    g = cb.clone_id_Gibbs(sp.csc_matrix(A2), sp.csc_matrix(np.nan_to_num(D)), Z, **kw)
    g0 = cb.clone_id_Gibbs(A, D, Z, **kw)
    assert np.array_equal(g["prob"], g0["prob"]) and g["logLik"][-1] == g0["logLik"][-1]
    # prior1 as a matrix: n_element rows
    N = A.shape[0]
    ok = cb.clone_id_Gibbs(A, D, Z, prior1=np.tile([0.45, 0.55], (N, 1)), **kw)
    assert np.array_equal(ok["prob"], g0["prob"])
    with pytest.raises(ValueError, match="prior1 need to be a matrix of n_element x 2"):
        cb.clone_id_Gibbs(A, D, Z, prior1=np.tile([0.45, 0.55], (N + 1, 1)), **kw)
    with pytest.raises(ValueError, match="prior1"):
        cb.clone_id_Gibbs(A, D, Z, prior1=np.tile([0.45, 0.55], (2, 1)), wise="global", **kw)
    with pytest.raises(cb.CdlError, match="one row per theta1 element"):        # the library's own check (R shim path)
        gpu_handle.gibbs_run(Z, None, min_iter=5, max_iter=5, prior1_matrix=np.tile([0.45, 0.55], (N + 2, 1)))
    off = cb.clone_id_Gibbs(A, D, Z, relax_Config=0, **kw)
    assert np.array_equal(off["Config_prob"], Z)                                # 0 is FALSE: Config never flips
