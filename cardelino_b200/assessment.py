"""Consumers of the `prob` matrix that clone_id() returns -- host-side mirror of R/assessment.R.

SURVEY.md section 8(f) item 4: `assign_scores` and the precision / recall helpers it is built from.  The reference
walks the cutoffs in an R loop and tests every cell at each one (O(cutoffs x cells), `R/assessment.R:145-166,
214-233,311-318`); at 10^6 cells that loop is what a user waits for after the sampler has finished, so here each
curve is one sort plus prefix sums (O(cells log cells + cutoffs log cells)) with the same outputs: same names,
arguments, defaults and returned fields (`df` is a dict of equally long NumPy arrays, a pandas DataFrame when
pandas is importable and `as_frame=True`).  Column indices are 0-based, like `colMatch` in clone_id.py.
"""
import numpy as np

from .clone_id import colMatch

__all__ = ["rowMax", "get_prob_value", "rowArgmax", "get_prob_label", "binaryPRC", "binaryROC", "multiPRC",
           "assign_scores"]


def rowMax(X, mode="best"):
    """R/assessment.R:36-49 -- per row: the highest value (`best`), the second highest (`second`) or their
    difference (`delta`)."""
    X = np.asarray(X, dtype=np.float64)
    if mode in ("second", "delta"):
        if X.shape[1] < 2:
            return np.full(X.shape[0], np.nan)                  # sorted_val[2] is NA in the reference
        top2 = -np.partition(-X, 1, axis=1)[:, :2]
        return top2[:, 1] if mode == "second" else top2[:, 0] - top2[:, 1]
    return X.max(axis=1)


get_prob_value = rowMax


def rowArgmax(X):
    """R/assessment.R:64-70 -- column of the row maximum, the earliest one on ties (0-based)."""
    return np.asarray(X).argmax(axis=1)


get_prob_label = rowArgmax


def _frame(cols, as_frame):
    if as_frame:
        try:
            import pandas as pd
            return pd.DataFrame(cols)
        except ImportError:
            pass
    return cols


def _selected(scores, cutoff, cut_direction):
    """For every cutoff: how many of the ascending-sorted scores pass it, and whether they are the head or the
    tail of that order."""
    s = np.sort(scores)
    n = s.size
    if cut_direction == "<=":
        return np.searchsorted(s, cutoff, side="right"), True
    if cut_direction == "<":
        return np.searchsorted(s, cutoff, side="left"), True
    if cut_direction == ">":
        return n - np.searchsorted(s, cutoff, side="right"), False
    return n - np.searchsorted(s, cutoff, side="left"), False          # ">=", and anything else, as in the reference


def _count_in_selection(scores, flags, count, head):
    """Number of set `flags` among the `count` smallest (head) or largest scores, for an array of counts."""
    order = np.argsort(scores, kind="stable")
    pre = np.concatenate(([0], np.cumsum(np.asarray(flags, dtype=np.int64)[order])))
    return pre[count] if head else pre[-1] - pre[scores.size - count]


def _trapezoid(x, y):
    """sum (x_i - x_{i+1}) (y_i + y_{i+1}) / 2, normalised by x_1 - x_last (R/assessment.R:177-182)."""
    with np.errstate(invalid="ignore", divide="ignore"):
        area = np.sum((x[:-1] - x[1:]) * (y[:-1] + y[1:]) * 0.5)
        return area / (x[0] - x[-1])


def binaryPRC(scores, labels, cutoff=None, cut_direction=">=", add_cut1=False, empty_precision=1, as_frame=False):
    """R/assessment.R:145-189 -- precision / recall of `scores` against binary `labels` at every cutoff."""
    scores = np.asarray(scores, dtype=np.float64).ravel()
    truth = np.asarray(labels).ravel() != 0                     # `labels == 1 | labels`
    cutoff = np.unique(scores) if cutoff is None else np.asarray(cutoff, dtype=np.float64).ravel()
    count, head = _selected(scores, cutoff, cut_direction)
    hit = _count_in_selection(scores, truth, count, head)
    with np.errstate(invalid="ignore", divide="ignore"):
        recall = hit / truth.sum()
        precision = hit / count                                 # mean of an empty selection is NaN, as in R
    if empty_precision is not None:
        precision = np.where(recall == 0, float(empty_precision), precision)
    if add_cut1:
        cutoff, recall, precision = np.append(cutoff, 1.0), np.append(recall, 0.0), np.append(precision, 1.0)
    return {"df": _frame({"cutoff": cutoff, "Recall": recall, "Precision": precision}, as_frame),
            "AUC": float(_trapezoid(recall, precision))}


def binaryROC(scores, labels, cutoff=None, cut_direction=">=", add_cut1=True, cutoff_point=0.9, as_frame=False):
    """R/assessment.R:214-262 -- ROC and precision / recall areas; the cutoffs always include 0, 1 and
    `cutoff_point`."""
    scores = np.asarray(scores, dtype=np.float64).ravel()
    labels = np.asarray(labels).ravel()
    cutoff = np.unique(scores) if cutoff is None else np.asarray(cutoff, dtype=np.float64).ravel()
    cutoff = np.unique(np.concatenate((cutoff, [cutoff_point, 0.0, 1.0])))
    pos, neg = labels == 1, labels == 0
    count, head = _selected(scores, cutoff, cut_direction)
    tp = _count_in_selection(scores, pos, count, head)
    fp = _count_in_selection(scores, neg, count, head)
    with np.errstate(invalid="ignore", divide="ignore"):
        fpr, tpr = fp / neg.sum(), tp / pos.sum()
        precision = tp / count
    precision = np.where((tpr == 0) & np.isnan(precision), 1.0, precision)
    if add_cut1:
        cutoff, fpr = np.append(cutoff, 1.0), np.append(fpr, 0.0)
        tpr, precision = np.append(tpr, 0.0), np.append(precision, 1.0)
    return {"df": _frame({"cutoff": cutoff, "TPR": tpr, "FPR": fpr, "Precision": precision}, as_frame),
            "AUC": float(_trapezoid(fpr, tpr)), "AUPRC": float(_trapezoid(tpr, precision))}


def multiPRC(prob_mat, simu_mat, marginal_mode="best", cutoff=None, multiLabel_rm=True, add_cut1=False,
             as_frame=False, verbose=True):
    """R/assessment.R:281-335 -- precision (argmax agrees with the simulated identity) and recall (share of
    cells whose score passes the cutoff) of a multi-class assignment.  `multiLabel_rm` is the reference's
    `multiLabel.rm`."""
    prob_mat = np.asarray(prob_mat, dtype=np.float64)
    simu_mat = np.asarray(simu_mat, dtype=np.float64)
    if prob_mat.shape != simu_mat.shape:
        raise ValueError("The shapes are not matched: prob_mat, simu_mat.")
    multi = (simu_mat > 0).sum(axis=1) > 1
    if multi.any() and verbose:
        print("%d samples have multiple labels." % int(multi.sum()))
    assign_0 = rowArgmax(simu_mat)
    assign_1 = rowArgmax(prob_mat)
    assign_0 = np.where(multi, -1, assign_0)
    prob_val = prob_mat.sum(axis=1) if marginal_mode == "sum" else rowMax(prob_mat, mode=marginal_mode)
    if multiLabel_rm:
        keep = ~multi
        prob_val, assign_0, assign_1 = prob_val[keep], assign_0[keep], assign_1[keep]
    cutoff = np.unique(prob_val) if cutoff is None else np.asarray(cutoff, dtype=np.float64).ravel()
    count, head = _selected(prob_val, cutoff, ">=")
    right = _count_in_selection(prob_val, assign_0 == assign_1, count, head)
    with np.errstate(invalid="ignore", divide="ignore"):
        recall = count / prob_val.size
        precision = right / count
    if add_cut1:
        cutoff, recall, precision = np.append(cutoff, 1.0), np.append(recall, 0.0), np.append(precision, 1.0)
    return {"df": _frame({"cutoff": cutoff, "Recall": recall, "Precision": precision}, as_frame),
            "AUC": float(_trapezoid(recall, precision))}


def assign_scores(prob, I_sim, cutoff=None, as_frame=False, verbose=True):
    """R/assessment.R:354-373 -- score a simulated donor: singlet assignment (multiPRC on the clone columns matched
    to the truth by colMatch) and doublet detection (binaryPRC on the row sums, direction `<=`)."""
    prob = np.asarray(prob, dtype=np.float64)
    I_sim = np.asarray(I_sim, dtype=np.float64)
    if cutoff is None:
        cutoff = np.minimum(np.arange(0, 1001) * 0.001, 1.0)    # seq(0, 1, 0.001) = from + (0:n) * by, capped at `to`
    col_idx = colMatch(I_sim, prob)
    if verbose:
        print(col_idx)
    sg = multiPRC(prob[:, col_idx], I_sim, multiLabel_rm=True, cutoff=cutoff, as_frame=as_frame, verbose=verbose)
    db = binaryPRC(prob.sum(axis=1), (I_sim > 0).sum(axis=1) > 1, cut_direction="<=", cutoff=cutoff,
                   as_frame=as_frame)
    return {"df_sg": sg["df"], "df_db": db["df"], "AUC_sg": sg["AUC"], "AUC_db": db["AUC"]}
