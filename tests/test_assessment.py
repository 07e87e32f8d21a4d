"""cardelino_b200/assessment.py (sort + prefix sums) against the loop-for-loop restatement of R/assessment.R in
oracle/assessment_oracle.py, and the worked example of the reference's documentation."""
import numpy as np
import pytest

import cardelino_b200 as cb
from oracle import assessment_oracle as AO


def _same(got, ref, keys):
    for k in keys:
        assert np.allclose(np.asarray(got["df"][k], dtype=float), ref[k], rtol=1e-12, atol=0, equal_nan=True), k
    for k in ("AUC", "AUPRC"):
        if k in ref:
            assert (np.isnan(got[k]) and np.isnan(ref[k])) or got[k] == pytest.approx(ref[k], rel=1e-12), k


def test_doc_example_binary_prc():
    """R/assessment.R:136-142: scores 1:10 against 0 0 0 1 0 1 0 1 1 1."""
    scores = np.arange(1, 11)
    labels = np.array([0, 0, 0, 1, 0, 1, 0, 1, 1, 1])
    r = cb.binaryPRC(scores, labels)
    assert np.array_equal(r["df"]["cutoff"], np.arange(1, 11))
    assert np.allclose(r["df"]["Recall"], [1, 1, 1, 1, .8, .8, .6, .6, .4, .2])
    assert np.allclose(r["df"]["Precision"], [.5, 5 / 9, 5 / 8, 5 / 7, 4 / 6, 4 / 5, 3 / 4, 1, 1, 1])
    for kw in (dict(cutoff=np.arange(1, 10, 2)), dict(cut_direction=">"), dict(add_cut1=True),
               dict(cut_direction="<="), dict(cut_direction="<", empty_precision=None)):
        _same(cb.binaryPRC(scores, labels, **kw), AO.binaryPRC(scores, labels, **kw), ("cutoff", "Recall", "Precision"))
        _same(cb.binaryROC(scores / 10, labels, **{k: v for k, v in kw.items() if k != "empty_precision"}),
              AO.binaryROC(scores / 10, labels, **{k: v for k, v in kw.items() if k != "empty_precision"}),
              ("cutoff", "TPR", "FPR", "Precision"))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_curves_match_the_loops_on_random_assignments(seed):
    rng = np.random.default_rng(seed)
    M, K = 600, 5
    truth = rng.integers(0, K, M)
    I_sim = np.zeros((M, K))
    I_sim[np.arange(M), truth] = 1
    doublet = rng.random(M) < 0.1
    I_sim[doublet, (truth[doublet] + 1) % K] = 1                       # two labels: a simulated doublet
    logits = rng.normal(size=(M, K)) + 3.0 * I_sim
    prob = np.exp(logits) / np.exp(logits).sum(1, keepdims=True)
    prob = np.round(prob, 2)                                           # ties between cells and with the cutoffs
    prob[doublet] *= 0.7                                               # doublets: lower row sums
    sure = (rng.random(M) < 0.2) & ~doublet                            # saturated cells: prob exactly 1, as real
    prob[sure] = I_sim[sure]                                           # posteriors have (no empty selection at cutoff 1)
    assert np.array_equal(cb.rowArgmax(prob), AO.rowArgmax(prob))
    for mode in ("best", "second", "delta"):
        assert np.array_equal(cb.rowMax(prob, mode), AO.rowMax(prob, mode))
    cut = np.minimum(np.arange(0, 1001) * 0.001, 1.0)
    for kw in (dict(), dict(cutoff=cut), dict(marginal_mode="delta", add_cut1=True), dict(multiLabel_rm=False),
               dict(marginal_mode="sum", cutoff=cut)):
        _same(cb.multiPRC(prob, I_sim, verbose=False, **kw), AO.multiPRC(prob, I_sim, **kw),
              ("cutoff", "Recall", "Precision"))
    scores, labels = prob.sum(1), doublet
    for d in (">=", ">", "<=", "<"):
        _same(cb.binaryPRC(scores, labels, cut_direction=d, cutoff=cut), AO.binaryPRC(scores, labels, cut_direction=d, cutoff=cut),
              ("cutoff", "Recall", "Precision"))
        _same(cb.binaryROC(scores, labels.astype(int), cut_direction=d), AO.binaryROC(scores, labels.astype(int), cut_direction=d),
              ("cutoff", "TPR", "FPR", "Precision"))
    # assign_scores = colMatch + multiPRC on the matched columns + binaryPRC on the row sums (R/assessment.R:354-373)
    perm = rng.permutation(K)
    s = cb.assign_scores(prob[:, perm], I_sim, verbose=False)
    idx = cb.colMatch(I_sim, prob[:, perm])
    assert np.array_equal(np.asarray(perm)[idx], np.arange(K))         # the columns are matched back
    sg = AO.multiPRC(prob, I_sim, cutoff=cut)
    db = AO.binaryPRC(prob.sum(1), doublet, cut_direction="<=", cutoff=cut)
    assert np.isfinite(sg["AUC"]) and np.isfinite(db["AUC"])
    assert s["AUC_sg"] == pytest.approx(sg["AUC"], rel=1e-12) and s["AUC_db"] == pytest.approx(db["AUC"], rel=1e-12)
    assert np.allclose(s["df_sg"]["Precision"], sg["Precision"], equal_nan=True)
    assert 0.5 < s["AUC_sg"] <= 1.0


def test_shape_error_message_and_frames():
    with pytest.raises(ValueError, match="The shapes are not matched: prob_mat, simu_mat."):
        cb.multiPRC(np.zeros((3, 2)), np.zeros((3, 3)))
    df = cb.binaryPRC(np.arange(5.0), [0, 1, 0, 1, 1], as_frame=True)["df"]
    assert list(df.columns) == ["cutoff", "Recall", "Precision"] and len(df) == 5


def test_large_donor_scores_fast():
    """10^6 cells x 16 clones at the reference's 1001 cutoffs: one sort per curve, not 1001 passes."""
    import time
    rng = np.random.default_rng(5)
    M, K = 1_000_000, 16
    truth = rng.integers(0, K, M)
    I_sim = np.zeros((M, K), dtype=np.float32)
    I_sim[np.arange(M), truth] = 1
    prob = rng.random((M, K), dtype=np.float32) * 0.05
    prob[np.arange(M), np.where(rng.random(M) < 0.9, truth, (truth + 1) % K)] += 0.6
    prob /= prob.sum(1, keepdims=True)
    t0 = time.perf_counter()
    r = cb.multiPRC(prob, I_sim, cutoff=np.arange(0, 1001) * 0.001, verbose=False)
    dt = time.perf_counter() - t0
    assert 0.85 < r["df"]["Precision"][0] < 0.95 and dt < 20
