"""TEST INFRASTRUCTURE ONLY -- loop-for-loop restatement of the scoring helpers of R/assessment.R, the checker
for cardelino_b200/assessment.py (which computes the same curves with one sort and prefix sums).

PARITY UNPINNED like the rest of oracle/: R is not installed here and the reference's tests hold no numbers for
these functions; the restatement follows the R source line by line (0-based column indices).

    rowMax       R/assessment.R:36-49        rowArgmax    R/assessment.R:64-70
    binaryPRC    R/assessment.R:145-189      binaryROC    R/assessment.R:214-262
    multiPRC     R/assessment.R:281-335
"""
import numpy as np


def rowMax(X, mode="best"):
    X = np.asarray(X, dtype=np.float64)
    out = np.zeros(X.shape[0])
    for i in range(X.shape[0]):
        sv = np.sort(X[i])[::-1]
        if mode == "delta":
            out[i] = sv[0] - sv[1]
        elif mode == "second":
            out[i] = sv[1]
        else:
            out[i] = sv[0]
    return out


def rowArgmax(X):
    X = np.asarray(X)
    return np.array([int(np.argmax(X[j])) for j in range(X.shape[0])])


def _pick(scores, c, cut_direction):
    if cut_direction == "<=":
        return scores <= c
    if cut_direction == "<":
        return scores < c
    if cut_direction == ">":
        return scores > c
    return scores >= c


def _auc(x, y):
    auc = 0.0
    for i in range(len(x) - 1):
        auc = (x[i] - x[i + 1]) * (y[i] + y[i + 1]) * 0.5 + auc
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.float64(auc) / np.float64(x[0] - x[-1])


def binaryPRC(scores, labels, cutoff=None, cut_direction=">=", add_cut1=False, empty_precision=1):
    scores = np.asarray(scores, dtype=np.float64)
    truth = np.asarray(labels) != 0
    if cutoff is None:
        cutoff = np.sort(np.unique(scores))
    cutoff = list(np.asarray(cutoff, dtype=np.float64))
    n_positive = truth.sum()
    P, R = [], []
    for c in cutoff:
        idx = _pick(scores, c, cut_direction)
        t = truth[idx]
        with np.errstate(invalid="ignore", divide="ignore"):
            R.append(np.float64(t.sum()) / np.float64(n_positive))
            P.append(np.float64(t.sum()) / np.float64(t.size))
    P, R = np.array(P), np.array(R)
    if empty_precision is not None:
        P[R == 0] = empty_precision
    if add_cut1:
        cutoff, R, P = cutoff + [1.0], np.append(R, 0.0), np.append(P, 1.0)
    return {"cutoff": np.array(cutoff), "Recall": R, "Precision": P, "AUC": float(_auc(R, P))}


def binaryROC(scores, labels, cutoff=None, cut_direction=">=", add_cut1=True, cutoff_point=0.9):
    scores = np.asarray(scores, dtype=np.float64)
    labels = np.asarray(labels)
    if cutoff is None:
        cutoff = np.sort(np.unique(scores))
    cutoff = list(np.sort(np.unique(np.concatenate((np.asarray(cutoff, dtype=np.float64), [cutoff_point, 0, 1])))))
    T, F, P = [], [], []
    for c in cutoff:
        idx = _pick(scores, c, cut_direction)
        with np.errstate(invalid="ignore", divide="ignore"):
            F.append(np.float64((labels[idx] == 0).sum()) / np.float64((labels == 0).sum()))
            T.append(np.float64((labels[idx] == 1).sum()) / np.float64((labels == 1).sum()))
            P.append(np.float64((labels[idx] == 1).sum()) / np.float64(idx.sum()))
    T, F, P = np.array(T), np.array(F), np.array(P)
    P[(T == 0) & np.isnan(P)] = 1.0
    if add_cut1:
        cutoff, F, T, P = cutoff + [1.0], np.append(F, 0.0), np.append(T, 0.0), np.append(P, 1.0)
    return {"cutoff": np.array(cutoff), "TPR": T, "FPR": F, "Precision": P, "AUC": float(_auc(F, T)),
            "AUPRC": float(_auc(T, P))}


def multiPRC(prob_mat, simu_mat, marginal_mode="best", cutoff=None, multiLabel_rm=True, add_cut1=False):
    prob_mat = np.asarray(prob_mat, dtype=np.float64)
    simu_mat = np.asarray(simu_mat, dtype=np.float64)
    idx_multi = (simu_mat > 0).sum(axis=1) > 1
    a0 = rowArgmax(simu_mat)
    a1 = rowArgmax(prob_mat)
    a0[idx_multi] = -1
    pv = prob_mat.sum(axis=1) if marginal_mode == "sum" else rowMax(prob_mat, marginal_mode)
    if multiLabel_rm:
        keep = a0 >= 0
        pv, a1, a0 = pv[keep], a1[keep], a0[keep]
    if cutoff is None:
        cutoff = np.sort(np.unique(pv))
    cutoff = list(np.asarray(cutoff, dtype=np.float64))
    P, R = [], []
    for c in cutoff:
        idx = pv >= c
        R.append(idx.mean())
        with np.errstate(invalid="ignore", divide="ignore"):
            P.append(np.float64((a0 == a1)[idx].sum()) / np.float64(idx.sum()))
    P, R = np.array(P), np.array(R)
    if add_cut1:
        cutoff, R, P = cutoff + [1.0], np.append(R, 0.0), np.append(P, 1.0)
    return {"cutoff": np.array(cutoff), "Recall": R, "Precision": P, "AUC": float(_auc(R, P))}
