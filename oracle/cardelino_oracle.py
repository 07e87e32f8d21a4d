"""TEST INFRASTRUCTURE ONLY -- CPU restatement (float64 NumPy) of cardelino's clone-assignment path.

This file is the *checker* for the CUDA implementation.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it; the product path
(cardelino_b200/) never does and fails loudly when the CUDA library is missing.

PARITY UNPINNED: the reference is pure R, R is not installed in the build image, and the reference's
own tests (tests/testthat/test-clone_id.R:7-49) pin only types, never numbers.  Every function below
is restated from the R source it cites and from base-R semantics (nmath `sample`, `rbinom`, `lchoose`)
that live outside /root/reference.  tools/record_reference.R lets a user with R pin it for real.

Restated functions (reference file:line):
    preprocess        R/clone_id.R:376-391 (Gibbs), :256-270 (EM)
    get_logLik        R/clone_id.R:734-752
    geweke_z          R/clone_id.R:688-715
    devianceIC        R/clone_id.R:765-798
    clone_id_EM       R/clone_id.R:240-326
    clone_id_Gibbs    R/clone_id.R:351-675      (dense, op for op, injectable draws)
    clone_id          R/clone_id.R:78-186       (dispatcher + chain merge)
    colMatch          R/assessment.R:87-121
    r_sample_one      base R src/main/random.c ProbSampleReplace + src/main/sort.c revsort
    r_rbinom1         base R src/nmath/rbinom.c (size = 1 inversion branch)
"""
import itertools
import math

import numpy as np
from scipy.special import gammaln

from . import philox

# --------------------------------------------------------------------------------------------
# base-R semantics
# --------------------------------------------------------------------------------------------


def revsort(a, ib):
    """R's revsort (src/main/sort.c): heapsort `a` into DESCENDING order, permuting `ib` alongside.
    Not stable; tie order is whatever this exact sift-down produces, so it is restated literally
    (1-based indices as in the C source)."""
    n = len(a)
    if n <= 1:
        return
    a = a  # modified in place (python lists)
    l = (n >> 1) + 1
    ir = n
    while True:
        if l > 1:
            l -= 1
            ra = a[l - 1]
            ii = ib[l - 1]
        else:
            ra = a[ir - 1]
            ii = ib[ir - 1]
            a[ir - 1] = a[0]
            ib[ir - 1] = ib[0]
            ir -= 1
            if ir == 1:
                a[0] = ra
                ib[0] = ii
                return
        i = l
        j = l << 1
        while j <= ir:
            if j < ir and a[j - 1] > a[j]:
                j += 1
            if ra > a[j - 1]:
                a[i - 1] = a[j - 1]
                ib[i - 1] = ib[j - 1]
                i = j
                j += i
            else:
                j = ir + 1
        a[i - 1] = ra
        ib[i - 1] = ii


def r_sample_one(prob, u, return_cdf=False):
    """`sample(seq_len(K), 1, replace=TRUE, prob=prob)` with the uniform `u` injected
    (R/clone_id.R:493).  FixupProb normalises; ProbSampleReplace sorts p descending (revsort),
    cumulates, returns the first j with u <= cum[j] (the last one if none).  Returns 1-based label."""
    p = [float(x) for x in prob]
    s = 0.0
    for x in p:
        s += x
    p = [x / s for x in p]
    perm = list(range(1, len(p) + 1))
    revsort(p, perm)
    for i in range(1, len(p)):
        p[i] += p[i - 1]
    j = 0
    nm1 = len(p) - 1
    while j < nm1:
        if u <= p[j]:
            break
        j += 1
    if return_cdf:
        return perm[j], p, perm
    return perm[j]


def r_rbinom1(p, u):
    """`rbinom(1, size=1, prob=p)` with its uniform injected (R/clone_id.R:534).  p == 0 / p == 1
    return without consuming a uniform; otherwise inversion on q = 1 - min(p, 1-p), mirrored when
    p > 0.5.  Vectorised over arrays."""
    p = np.asarray(p, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    pm = np.minimum(p, 1.0 - p)
    q = 1.0 - pm
    ix = np.where(u < q, 0, 1)
    ix = np.where(p > 0.5, 1 - ix, ix)
    ix = np.where(p == 0.0, 0, ix)
    ix = np.where(p == 1.0, 1, ix)
    return ix.astype(np.float64)


def lchoose(n, k):
    return gammaln(n + 1.0) - gammaln(k + 1.0) - gammaln(n - k + 1.0)


# --------------------------------------------------------------------------------------------
# deterministic pieces
# --------------------------------------------------------------------------------------------


def preprocess(A, D):
    """R/clone_id.R:380-391.  A, D float arrays (variants x cells), NaN = NA.  Returns A1, B1 (NA->0)
    and W_log = sum(lchoose(D, A), na.rm=TRUE)."""
    A = np.array(A, dtype=np.float64, copy=True)
    D = np.array(D, dtype=np.float64, copy=True)
    with np.errstate(invalid="ignore"):
        zero = D == 0                      # which() drops NA
        A[zero] = np.nan
        D[zero] = np.nan
        A[(D > 0) & np.isnan(A)] = 0.0     # NA in a logical subscript is skipped on assignment
    with np.errstate(invalid="ignore"):
        B = D - A
    ok = ~np.isnan(D) & ~np.isnan(A)
    W_log = float(np.sum(lchoose(D[ok], A[ok])))
    A1 = np.nan_to_num(A, nan=0.0)
    B1 = np.nan_to_num(B, nan=0.0)
    return A1, B1, W_log


def get_logLik(A1, B1, Config, Assign, theta0, theta1):
    """R/clone_id.R:734-752, dense, including the log(exp(.)) form (quirk: underflows to -Inf for
    very deep entries).  Assign is a 1-based label vector or an M x K probability matrix."""
    Config = np.asarray(Config, dtype=np.float64)
    Assign = np.asarray(Assign)
    if Assign.ndim == 1:
        P = np.zeros((len(Assign), Config.shape[1]))
        P[np.arange(len(Assign)), Assign.astype(int) - 1] = 1.0
    else:
        P = Assign.astype(np.float64)
    prob_mat = Config @ P.T
    theta1 = np.asarray(theta1, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        Lik = (np.exp(np.log(theta1) * A1 + np.log(1 - theta1) * B1) * prob_mat +
               np.exp(np.log(theta0) * A1 + np.log(1 - theta0) * B1) * (1 - prob_mat))
        ll = np.log(Lik)
    return float(np.nansum(ll) + np.sum(lchoose(A1 + B1, A1)))


def get_logLik_tables(A1, B1, Config, assign, theta0, theta1_vec):
    """Same quantity for a label vector through the N x K sufficient statistics:
    finite where the dense log(exp(.)) form underflows.  theta1_vec has one value per variant."""
    N, M = A1.shape
    K = Config.shape[1]
    Z = np.zeros((M, K))
    Z[np.arange(M), np.asarray(assign, int) - 1] = 1.0
    SA = A1 @ Z
    SB = B1 @ Z
    t1 = np.broadcast_to(np.asarray(theta1_vec, dtype=np.float64).reshape(-1, 1), (N, 1))
    ll = (Config * (SA * np.log(t1) + SB * np.log(1 - t1)) +
          (1 - Config) * (SA * math.log(theta0) + SB * math.log(1 - theta0)))
    return float(ll.sum() + np.sum(lchoose(A1 + B1, A1)))


def geweke_z(X, first=0.1, last=0.5):
    """R/clone_id.R:688-715.  Both window variances divide by nrow(A) - 1 (reference quirk)."""
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 1:
        X = X.reshape(-1, 1)
    n = X.shape[0]
    Aw = X[: int(math.floor(first * n))]
    Bw = X[int(math.ceil(last * n)) - 1:]
    if Aw.shape[0] == 0:                       # mean of an empty window is NaN in R as well
        return np.full(X.shape[1], np.nan)
    with np.errstate(invalid="ignore", divide="ignore"):
        mA = Aw.mean(axis=0)
        mB = Bw.mean(axis=0)
        vA = ((Aw - mA) ** 2).sum(axis=0) / (Aw.shape[0] - 1)
        vB = ((Bw - mB) ** 2).sum(axis=0) / (Aw.shape[0] - 1)
        return (mA - mB) / np.sqrt(vA + vB + 1e-50)


def geweke_z_segment_sums(X, first=0.1, last=0.5, storage=np.float64):
    """The same Z from column sums of x and x^2 over SEGMENTS of rows, as the streaming mode of the CUDA path evaluates
    it (csrc/api.cu, csrc/post.cu::k_geweke_sums): the rows are cut at floor(first n) and ceil(last n) - 1, each
    window is the SUM of its segments (kept in `storage` precision: fp64, or fp32 like the default snapshots), and
    the variances are one-pass, clamped at 0.  Nothing is ever obtained as a difference of two prefix sums."""
    X = np.asarray(X, dtype=np.float64)
    n = X.shape[0]
    nA = math.floor(first * n)
    rb = math.ceil(last * n) - 1
    nB = n - rb
    seg = lambda lo, hi: (X[lo:hi].sum(axis=0).astype(storage).astype(np.float64),
                          (X[lo:hi] ** 2).sum(axis=0).astype(storage).astype(np.float64))
    a1, a2 = seg(0, nA)
    b1, b2 = seg(rb, n)
    with np.errstate(invalid="ignore", divide="ignore"):
        mA, mB = a1 / nA, b1 / nB
        vA = np.maximum(a2 - nA * mA * mA, 0.0) / (nA - 1)
        vB = np.maximum(b2 - nB * mB * mB, 0.0) / (nA - 1)
        return (mA - mB) / np.sqrt(vA + vB + 1e-50)


def geweke_z_prefix_differences(X, first=0.1, last=0.5):
    """Round 1's form, kept as the counter-example: window B as a DIFFERENCE of prefix sums.  A column that held large
    values in rows before the window (a cell that visited the clone early) and tiny ones inside it loses the
    window's sums to cancellation; tests/test_oracle.py shows the verdicts that flip."""
    X = np.asarray(X, dtype=np.float64)
    n = X.shape[0]
    s1, s2 = np.cumsum(X, axis=0), np.cumsum(X * X, axis=0)
    zero = np.zeros(X.shape[1])
    at = lambda S, r: S[r - 1] if r >= 1 else zero
    nA = math.floor(first * n)
    rb = math.ceil(last * n) - 1
    nB = n - rb
    with np.errstate(invalid="ignore", divide="ignore"):
        mA = at(s1, nA) / nA
        b1, b2 = s1[-1] - at(s1, rb), s2[-1] - at(s2, rb)
        mB = b1 / nB
        vA = np.maximum(at(s2, nA) - nA * mA * mA, 0.0) / (nA - 1)
        vB = np.maximum(b2 - nB * mB * mB, 0.0) / (nA - 1)
        return (mA - mB) / np.sqrt(vA + vB + 1e-50)


def r_round(x, digits=0):
    """R's round() to `digits` decimals (half-even on the decimal representation is not needed at the
    3-digit precision the sampler prints)."""
    f = 10.0 ** digits
    return np.round(x * f) / f


def converged_fraction(X):
    """mean(abs(Geweke_Z(X)) <= 2, na.rm=TRUE), R/clone_id.R:581-589."""
    Z = geweke_z(X)
    ok = ~np.isnan(Z)
    if not ok.any():
        return float("nan")
    return float(np.mean(np.abs(Z[ok]) <= 2))


def devianceIC(logLik_all, logLik_post):
    """R/clone_id.R:765-798."""
    logLik_all = np.asarray(logLik_all, dtype=np.float64)
    m = float(np.mean(logLik_all))
    v = float(np.var(logLik_all, ddof=1)) if logLik_all.size > 1 else float("nan")
    p_D_S = -2 * m - (-2 * logLik_post)
    DIC_S = -2 * logLik_post + 2 * p_D_S
    p_D_G = 2 * v
    DIC_G = -2 * logLik_post + 2 * p_D_G
    return {"DIC": DIC_G, "logLik_var": v, "D_mean": -2 * m, "D_post": -2 * logLik_post,
            "DIC_Gelman": DIC_G, "DIC_Spiegelhalter": DIC_S}


def colMatch(A, B, force=False):
    """R/assessment.R:87-121.  Returns 0-based column indices of B matched to A's columns."""
    A = np.asarray(A, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    if A.shape[0] != B.shape[0]:
        raise ValueError("A and B have different rows.")
    if A.shape[1] > B.shape[1]:
        raise ValueError("A must have equal or smaller columns than B.")
    ka, kb = A.shape[1], B.shape[1]
    if force:
        best, best_idx = None, None
        for perm in permn(kb):
            idx = list(perm[:ka])
            mae = np.mean(np.abs(A - B[:, idx]))
            if best is None or mae < best:      # which.min keeps the first minimum
                best, best_idx = mae, idx
        return np.array(best_idx)
    MAE = np.zeros((ka, kb))
    for i in range(ka):
        for j in range(kb):
            MAE[i, j] = np.mean(np.abs(A[:, i] - B[:, j]))
    tmp = MAE.copy()
    out = -np.ones(ka, dtype=int)
    while (out == -1).any():
        # which(x == min(x), arr.ind=TRUE) lists matches in column-major order; the reference takes
        # idx_tmp[1], idx_tmp[2] of that matrix: with one match these are (row, col); with several
        # matches idx_tmp[2] is the *row of the second match* (reference quirk, not reproduced: ties
        # in mean-abs-difference of real-valued posteriors do not occur in practice).
        flat = np.argmin(tmp.T.ravel())
        c, r = divmod(flat, ka)
        out[r] = c
        tmp[r, :] = MAE.max() + 1
        tmp[:, c] = MAE.max() + 1
    return out


def permn(n):
    """combinat::permn(n) enumeration order (Johnson-Trotter style adjacent transpositions as in
    combinat 0.0-8).  Only the argmin over all permutations matters to colMatch(force=TRUE); the order
    decides ties only, so plain lexicographic order is used and documented as such."""
    return itertools.permutations(range(n))


# --------------------------------------------------------------------------------------------
# draw sources (the reference draws from R's global Mersenne-Twister stream; here draws are injected)
# --------------------------------------------------------------------------------------------


class NumpyDraws:
    """Free-running chain: NumPy generator for every draw (statistical comparisons only)."""

    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)

    def beta(self, a, b, what, it):
        return self.rng.beta(a, b)

    def u_assign(self, it, M):
        return self.rng.random(M)

    def u_config(self, it, NK):
        return self.rng.random(NK)

    def runif(self, lo, hi):
        return self.rng.uniform(lo, hi)


class ReplayDraws:
    """Teacher-forced chain: categorical / Bernoulli uniforms come from Philox with the CUDA
    sampler's keying (or from recorded arrays), Beta draws come from a recorded run's traces."""

    def __init__(self, seed=0, chain=0, theta0_all=None, theta1_all=None, relax_rate_all=None,
                 u_assign=None, u_config=None):
        self.seed, self.chain = seed, chain
        self.theta0_all = None if theta0_all is None else np.asarray(theta0_all, float).reshape(-1)
        self.theta1_all = None if theta1_all is None else np.asarray(theta1_all, float)
        if self.theta1_all is not None and self.theta1_all.ndim == 1:
            self.theta1_all = self.theta1_all.reshape(-1, 1)
        self.relax_rate_all = None if relax_rate_all is None else np.asarray(relax_rate_all, float).reshape(-1)
        self._u_assign, self._u_config = u_assign, u_config

    def beta(self, a, b, what, it):
        # `it` is the 1-based row of the *_all trace being filled
        if what == "theta0":
            return self.theta0_all[it - 1]
        if what == "theta1":
            return self.theta1_all[it - 1]
        if what == "relax":
            return self.relax_rate_all[it - 1]
        raise KeyError(what)

    def u_assign(self, it, M):
        if self._u_assign is not None:
            return self._u_assign[it - 1]
        return philox.uniforms(self.seed, self.chain, it, philox.STREAM_ASSIGN, np.arange(M))

    def u_config(self, it, NK):
        if self._u_config is not None:
            return self._u_config[it - 1]
        return philox.uniforms(self.seed, self.chain, it, philox.STREAM_CONFIG, np.arange(NK))


# --------------------------------------------------------------------------------------------
# EM  (R/clone_id.R:240-326)
# --------------------------------------------------------------------------------------------


def clone_id_EM(A, D, Config, Psi=None, min_iter=10, max_iter=1000, logLik_threshold=1e-5,
                verbose=False, theta_init=None, draws=None):
    Config = np.asarray(Config, dtype=np.float64)
    K = Config.shape[1]
    if Psi is None:
        Psi = np.full(K, 1.0 / K)
    Psi = np.asarray(Psi, dtype=np.float64)
    A = np.asarray(A, float)
    D = np.asarray(D, float)
    if A.shape != D.shape or A.shape[0] != Config.shape[0] or K != len(Psi):
        raise ValueError("A and D must have the same size;\nA and Config must have the same number of "
                         "variants;\nConfig and Psi must have the same number of clones")
    A1, B1, W_log = preprocess(A, D)
    C1 = Config
    C0 = 1 - Config
    S1 = A1.T @ C0
    S2 = B1.T @ C0
    S3 = A1.T @ C1
    S4 = B1.T @ C1
    if theta_init is None:
        draws = draws or NumpyDraws(0)
        theta = np.array([draws.runif(0.001, 0.25), draws.runif(0.3, 0.6)])
    else:
        theta = np.array(theta_init, dtype=np.float64)
    logLik = 0.0
    logLik_new = logLik + logLik_threshold * 2
    it_done = 0
    prob_mat = None
    for it in range(1, max_iter + 1):
        it_done = it
        if it > min_iter and (logLik_new - logLik) < logLik_threshold:
            break
        logLik = logLik_new
        L = (S1 * math.log(theta[0]) + S2 * math.log(1 - theta[0]) +
             S3 * math.log(theta[1]) + S4 * math.log(1 - theta[1]))
        L = L + np.log(Psi)[None, :]
        mx = L.max(axis=1, keepdims=True)
        lse = mx[:, 0] + np.log(np.exp(L - mx).sum(axis=1))
        logLik_new = float(lse.sum() + W_log)
        E = np.exp(L - mx)
        prob_mat = E / E.sum(axis=1, keepdims=True)
        den0 = float(np.sum(prob_mat * (S1 + S2)))
        theta[0] = 0.02 if (np.isnan(den0) or den0 == 0) else float(np.sum(prob_mat * S1)) / den0
        den1 = float(np.sum(prob_mat * (S3 + S4)))
        theta[1] = 0.75 if (np.isnan(den1) or den1 == 0) else float(np.sum(prob_mat * S3)) / den1
    return {"theta": theta, "prob": prob_mat, "logLik": logLik, "n_iter": it_done,
            "logLik_new": logLik_new}


# --------------------------------------------------------------------------------------------
# Gibbs  (R/clone_id.R:351-675), dense and op-for-op
# --------------------------------------------------------------------------------------------


def _config_prior_oddlog(Config, relax_rate, keep_base_clone):
    """R/clone_id.R:450-456 / :510-516."""
    prior = Config.copy()
    prior[Config == 1] = 1 - relax_rate
    prior[Config == 0] = relax_rate
    if keep_base_clone:
        prior[:, 0] = Config[:, 0]
    with np.errstate(divide="ignore"):
        return np.log(prior) - np.log(1 - prior)


def clone_id_Gibbs(A, D, Config, Psi=None, relax_Config=True, relax_rate_fixed=None,
                   relax_rate_prior=(1, 9), keep_base_clone=True, prior0=(0.2, 99.8),
                   prior1=(0.45, 0.55), min_iter=5000, max_iter=20000, buin_frac=0.5,
                   wise="variant", relabel=False, verbose=False, draws=None, keep_assign=True):
    """Dense restatement.  `draws` supplies every random number (NumpyDraws / ReplayDraws).

    Conscious resolutions of reference defects:
      * relax_Config=FALSE: the reference fails with "object 'Config_new' not found" (:575); here
        Config_new == Config and Config_all rows hold Config (intended semantics).
      * wise="element" with uncovered entries: the reference propagates NA through %*% (:526-528);
        here uncovered entries are skipped (theta1 there is irrelevant because A1 = B1 = 0).
    Everything else, including the Geweke denominator and the old-theta logLik trace, is as written.
    """
    draws = draws or NumpyDraws(0)
    Config = np.asarray(Config, dtype=np.float64)
    K = Config.shape[1]
    if Psi is None:
        Psi = np.full(K, 1.0 / K)
    Psi = np.asarray(Psi, dtype=np.float64)
    A = np.asarray(A, float)
    D = np.asarray(D, float)
    if A.shape != D.shape or A.shape[0] != Config.shape[0] or K != len(Psi):
        raise ValueError("A and D must have the same size;\n A and Config must have the same number of "
                         "variants;\nConfig and Psi must have the same number of clones")
    if wise not in ("element", "variant", "global"):
        raise ValueError("Input wise mode: %s, while only supporting: element, variant, global" % wise)
    N, M = A.shape
    A1, B1, W_log = preprocess(A, D)
    covered = (A1 + B1) > 0

    def s_lists(C):
        return ([A1 * (1 - C[:, [k]]) for k in range(K)], [B1 * (1 - C[:, [k]]) for k in range(K)],
                [A1 * C[:, [k]] for k in range(K)], [B1 * C[:, [k]] for k in range(K)])

    S1l, S2l, S3l, S4l = s_lists(Config)

    if wise == "global":
        n_element = 1
    elif wise == "variant":
        n_element = N
    else:
        el_idx = np.flatnonzero(covered.T.ravel())          # column-major which(A1 + B1 > 0)
        n_element = el_idx.size
    prior1 = np.asarray(prior1, dtype=np.float64)
    if prior1.ndim == 1 and prior1.size == 2:
        prior1 = np.tile(prior1, (n_element, 1))
    if prior1.ndim != 2:
        raise ValueError("prior1 need to be a matrix of n_element x 2")

    prob_all = np.zeros((max_iter, M * K))
    logLik_all = np.zeros(max_iter)
    assign_all = np.zeros((max_iter, M))
    theta0_all = np.zeros(max_iter)
    theta1_all = np.zeros((max_iter, n_element))
    Config_all = np.zeros((max_iter, N * K))
    relax_rate_all = np.zeros(max_iter)

    relax = relax_Config is not None and relax_Config is not False
    Config_new = Config.copy()
    if relax:
        if relax_rate_fixed is not None:
            if relax_rate_fixed > 1 or relax_rate_fixed < 0:
                raise ValueError("relax_rate_fixed needs to be NULL or in [0, 1].")
            relax_rate = relax_rate_fixed
            relax_rate_all[:] = relax_rate
        elif relax_rate_prior is not None:
            relax_rate = relax_rate_prior[0] / (relax_rate_prior[0] + relax_rate_prior[1])
        else:
            raise ValueError("Require value for either relax_Config or relax_prior.")
        oddlog = _config_prior_oddlog(Config, relax_rate, keep_base_clone)
    else:
        Config_all[:] = Config.T.ravel()      # resolution of quirk 1

    theta0_all[0] = draws.beta(prior0[0], prior0[1], "theta0", 1)
    theta1_all[0, :] = draws.beta(prior1[:, 0], prior1[:, 1], "theta1", 1)

    def theta1_matrix(vec):
        if wise == "global":
            return np.full((N, M), vec[0])
        if wise == "variant":
            return np.broadcast_to(vec.reshape(N, 1), (N, M)).copy()
        t = np.full(N * M, 0.5)               # value irrelevant where A1 = B1 = 0
        t[el_idx] = vec
        return t.reshape(M, N).T.copy()

    logPsi = np.log(Psi)
    it_final = max_iter
    for it in range(2, max_iter + 1):
        theta0 = theta0_all[it - 2]
        theta1 = theta1_matrix(theta1_all[it - 2])
        l0, l0c = math.log(theta0), math.log(1 - theta0)
        l1, l1c = np.log(theta1), np.log(1 - theta1)
        L = np.zeros((M, K))
        for k in range(K):
            L[:, k] = ((S1l[k] * l0).sum(0) + (S2l[k] * l0c).sum(0) +
                       (S3l[k] * l1).sum(0) + (S4l[k] * l1c).sum(0)) + logPsi[k]
        Lamp = L - L.max(axis=1, keepdims=True)
        E = np.exp(Lamp)
        prob_mat = E / E.sum(axis=1, keepdims=True)
        prob_all[it - 1] = prob_mat.T.ravel()

        u = draws.u_assign(it, M)
        for j in range(M):
            assign_all[it - 1, j] = r_sample_one(prob_mat[j], u[j])
        z = assign_all[it - 1].astype(int) - 1

        if relax:
            if it > (0.1 * min_iter + 5) and relax_rate_fixed is None:
                diff0 = np.sum((Config == Config_new)[:, 1:])
                diff1 = np.sum((Config != Config_new)[:, 1:])
                relax_rate = draws.beta(relax_rate_prior[0] + diff1, relax_rate_prior[1] + diff0,
                                        "relax", it)
                relax_rate_all[it - 1] = relax_rate
                oddlog = _config_prior_oddlog(Config, relax_rate, keep_base_clone)
            Iden = np.zeros((M, K))
            Iden[np.arange(M), z] = 1.0
            P0 = A1 * l0 + B1 * l0c + W_log
            P1 = A1 * l1 + B1 * l1c + W_log
            odd = P1 @ Iden - P0 @ Iden
            odd = odd + oddlog
            odd[odd > 50] = 50
            odd[odd < -50] = -50
            p = np.exp(odd) / (np.exp(odd) + 1)
            uc = draws.u_config(it, N * K)
            Config_new = r_rbinom1(p.T.ravel(), uc).reshape(K, N).T.copy()
            Config_all[it - 1] = Config_new.T.ravel()
            S1l, S2l, S3l, S4l = s_lists(Config_new)

        S1w = S2w = 0.0
        S3w = np.zeros((N, M))
        S4w = np.zeros((N, M))
        for k in range(K):
            idx = np.flatnonzero(z == k)
            S1w += S1l[k][:, idx].sum()
            S2w += S2l[k][:, idx].sum()
            S3w[:, idx] += S3l[k][:, idx]
            S4w[:, idx] += S4l[k][:, idx]
        if wise == "global":
            s3 = np.array([S3w.sum()])
            s4 = np.array([S4w.sum()])
        elif wise == "variant":
            s3 = S3w.sum(axis=1)
            s4 = S4w.sum(axis=1)
        else:
            s3 = S3w.T.ravel()[el_idx]
            s4 = S4w.T.ravel()[el_idx]
        theta0_all[it - 1] = draws.beta(prior0[0] + S1w, prior0[1] + S2w, "theta0", it)
        theta1_all[it - 1, :] = draws.beta(prior1[:, 0] + s3, prior1[:, 1] + s4, "theta1", it)

        logLik_all[it - 1] = get_logLik(A1, B1, Config_new, assign_all[it - 1], theta0, theta1)

        if it >= min_iter and it % 100 == 0:
            frac = converged_fraction(prob_all[:it])
            if verbose:
                print("%s%% converged." % (r_round(frac, 3) * 100))
            if frac > 0.995:
                it_final = it
                break
    it = it_final
    frac = converged_fraction(prob_all[:it])
    percent_converged = r_round(frac, 3) * 100

    n_buin = int(math.ceil(it * buin_frac))
    sl = slice(n_buin - 1, it)

    # prob_variant  (:606-622)
    from scipy.stats import binom
    if wise == "element":
        a = A1.T.ravel()[el_idx]
        d = a + B1.T.ravel()[el_idx]
    else:
        a = A1.T.ravel()
        d = a + B1.T.ravel()
    pdf1 = np.zeros_like(a)
    pdf0 = np.zeros_like(a)
    for i in range(n_buin, it + 1):
        if wise == "element":
            t1 = theta1_all[i - 1]
        elif wise == "variant":
            t1 = np.tile(theta1_all[i - 1], M)
        else:
            t1 = theta1_all[i - 1, 0]
        pdf1 += binom.pmf(a, d, t1)
        pdf0 += binom.pmf(a, d, theta0_all[i - 1])
    pv = np.full(N * M, np.nan)
    if wise == "element":
        pv[el_idx] = pdf1 / (pdf1 + pdf0)
    else:
        pv[:] = pdf1 / (pdf1 + pdf0)
    prob_variant = pv.reshape(M, N).T.copy()

    if relabel:
        col_idx_use = np.arange(K)
        for ii in range(n_buin, it + 1):
            mat1 = prob_all[ii - 2].reshape(K, M).T
            mat2 = prob_all[ii - 1].reshape(K, M).T
            idx = colMatch(mat1, mat2, force=(K <= 5))
            col_idx_use = col_idx_use[idx]
            prob_all[ii - 1] = mat2[:, col_idx_use].T.ravel()
            Config_all[ii - 1] = Config_all[ii - 1].reshape(K, N).T[:, col_idx_use].T.ravel()

    prob = prob_all[sl].mean(axis=0).reshape(K, M).T.copy()
    Config_prob = Config_all[sl].mean(axis=0).reshape(K, N).T.copy()
    theta0_mean = float(theta0_all[sl].mean())
    theta1_mean = theta1_matrix(theta1_all[sl].mean(axis=0))
    if wise == "element":
        tm = np.full(N * M, np.nan)
        tm[el_idx] = theta1_all[sl].mean(axis=0)
        theta1_out = tm.reshape(M, N).T.copy()
    else:
        theta1_out = theta1_mean
    logLik_post = get_logLik(A1, B1, Config_prob, prob, theta0_mean, theta1_mean)
    DIC = devianceIC(logLik_all[sl], logLik_post)
    return {
        "theta0": theta0_mean, "theta1": theta1_out,
        "theta0_all": theta0_all[:it].reshape(-1, 1), "theta1_all": theta1_all[:it],
        "element": (el_idx if wise == "element" else np.arange(N * M)),
        "logLik": logLik_all[:it], "prob_all": prob_all[:it], "prob": prob,
        "prob_variant": prob_variant, "relax_rate": float(relax_rate_all[sl].mean()),
        "Config_prob": Config_prob, "Config_all": Config_all[:it],
        "relax_rate_all": relax_rate_all[:it], "DIC": DIC,
        "assign_all": assign_all[:it], "n_iter": it, "percent_converged": percent_converged,
        "logLik_post": logLik_post, "W_log": W_log,
    }


def clone_id(A, D, Config=None, n_clone=None, Psi=None, relax_Config=True, relax_rate_fixed=None,
             inference="sampling", n_chain=1, verbose=False, draws_list=None, **kw):
    """R/clone_id.R:78-186 on already-matched matrices (dimnames handling lives in the host wrapper)."""
    if inference not in ("sampling", "EM"):
        raise ValueError("'arg' should be one of 'sampling', 'EM'")
    if Config is None and n_clone is None:
        raise ValueError("Config and n_clone can't be NULL together.")
    if Config is None:
        Config = np.zeros((np.asarray(D).shape[0], n_clone))
        relax_Config = True
        relax_rate_fixed = 0.5
    if inference == "EM":
        return clone_id_EM(A, D, Config, verbose=verbose, **kw)
    outs = []
    for c in range(n_chain):
        dr = draws_list[c] if draws_list is not None else NumpyDraws(1000 + c)
        outs.append(clone_id_Gibbs(A, D, Config, Psi=Psi, relax_Config=relax_Config,
                                   relax_rate_fixed=relax_rate_fixed, verbose=verbose, draws=dr, **kw))
    out = dict(outs[0])
    out["n_chain"] = 1
    for ii in range(1, n_chain):
        out["n_chain"] += 1
        idx = colMatch(out["prob"], outs[ii]["prob"], force=True)
        out["prob"] = out["prob"] + outs[ii]["prob"][:, idx]
        out["relax_rate"] = out["relax_rate"] + outs[ii]["relax_rate"]
        out["Config_prob"] = out["Config_prob"] + outs[ii]["Config_prob"][:, idx]
    if n_chain > 1:
        out["prob"] = out["prob"] / n_chain
        out["relax_rate"] = out["relax_rate"] / n_chain
        out["Config_prob"] = out["Config_prob"] / n_chain
    return out


# --------------------------------------------------------------------------------------------
# sparse / table formulation of ONE sweep (used to teacher-force big cases cheaply)
# --------------------------------------------------------------------------------------------


def sweep_tables(A1, B1, Config_cur, theta0, theta1_vec, Psi, u_assign):
    """prob_mat and labels of one sweep for wise in {global, variant} given the state of the previous
    iteration (R/clone_id.R:468-497).  A1/B1 may be scipy.sparse (variants x cells)."""
    th1 = np.asarray(theta1_vec, float).reshape(-1, 1)
    l0, l0c = math.log(theta0), math.log(1 - theta0)
    WA = Config_cur * np.log(th1) + (1 - Config_cur) * l0
    WB = Config_cur * np.log(1 - th1) + (1 - Config_cur) * l0c
    L = np.asarray(A1.T @ WA + B1.T @ WB) + np.log(np.asarray(Psi, float))[None, :]
    E = np.exp(L - L.max(axis=1, keepdims=True))
    prob = E / E.sum(axis=1, keepdims=True)
    labels = np.array([r_sample_one(prob[j], u_assign[j]) for j in range(prob.shape[0])])
    return prob, labels


# --------------------------------------------------------------------------------------------
# step-wise (teacher-forced) verification of a recorded chain
# --------------------------------------------------------------------------------------------


def verify_gibbs_trace(A, D, Config, rec, Psi=None, relax_Config=True, relax_rate_fixed=None,
                       relax_rate_prior=(1, 9), keep_base_clone=True, min_iter=5000, wise="variant",
                       seed=0, chain=0, u_assign=None, u_config=None, tie_rtol=1e-12, cdf_tol=1e-6,
                       dense_loglik=True):
    """Check every sweep of a recorded chain `rec` (dict with theta0_all, theta1_all, relax_rate_all,
    prob_all, assign_all, Config_all, logLik in the reference's layouts) against the reference
    formulas, each sweep teacher-forced from the RECORDED state of the previous one (SURVEY.md
    section 8c protocol 2/3):

      prob_all[t]   is a pure function of theta[t-1], Config_all[t-1]            (R/clone_id.R:468-481)
      assign_all[t] follows from prob_all[t] and the uniform of cell j          (:492-497)
      Config_all[t] follows from assign_all[t], theta[t-1], relax rate and u     (:501-535)
      logLik[t]     from Config_all[t], assign_all[t], theta[t-1]                (:574-577)

    Label disagreements are classified: `boundary` = the uniform lies within cdf_tol of a CDF
    boundary; `near_tie` = two probabilities agree to tie_rtol so their order in sample()'s sort is
    decided by rounding; anything else is `hard` (a real bug)."""
    Config = np.asarray(Config, dtype=np.float64)
    N, K = Config.shape
    A1, B1, W_log = preprocess(A, D)
    M = A1.shape[1]
    Psi = np.full(K, 1.0 / K) if Psi is None else np.asarray(Psi, float)
    theta0_all = np.asarray(rec["theta0_all"], float).reshape(-1)
    theta1_all = np.asarray(rec["theta1_all"], float)
    if theta1_all.ndim == 1:
        theta1_all = theta1_all.reshape(-1, 1)
    T = len(theta0_all)
    relax = relax_Config is not None and relax_Config is not False
    out = {"max_dprob": 0.0, "boundary": 0, "near_tie": 0, "hard": 0, "config_boundary": 0,
           "config_hard": 0, "max_dloglik_rel": 0.0, "first_hard": None, "n_sweeps": T - 1}
    Config_prev = Config.copy()
    if relax:
        relax_rate = relax_rate_fixed if relax_rate_fixed is not None else \
            relax_rate_prior[0] / (relax_rate_prior[0] + relax_rate_prior[1])
    for it in range(2, T + 1):
        theta0 = theta0_all[it - 2]
        th1 = theta1_all[it - 2]
        th1v = np.full(N, th1[0]) if wise == "global" else th1
        ua = u_assign[it - 1] if u_assign is not None else \
            philox.uniforms(seed, chain, it, philox.STREAM_ASSIGN, np.arange(M))
        # (a) prob
        l0, l0c = math.log(theta0), math.log(1 - theta0)
        l1, l1c = np.log(th1v).reshape(-1, 1), np.log(1 - th1v).reshape(-1, 1)
        WA = Config_prev * l1 + (1 - Config_prev) * l0
        WB = Config_prev * l1c + (1 - Config_prev) * l0c
        L = A1.T @ WA + B1.T @ WB + np.log(Psi)[None, :]
        E = np.exp(L - L.max(axis=1, keepdims=True))
        prob = E / E.sum(axis=1, keepdims=True)
        rec_prob = np.asarray(rec["prob_all"][it - 1]).reshape(K, M).T
        out["max_dprob"] = max(out["max_dprob"], float(np.abs(prob - rec_prob).max()))
        # (b) labels
        rec_lab = np.asarray(rec["assign_all"][it - 1]).astype(int)
        for j in range(M):
            lab, cdf, perm = r_sample_one(prob[j], ua[j], return_cdf=True)
            if lab != rec_lab[j]:
                ps = sorted(prob[j], reverse=True)
                gaps = [abs(ps[q] - ps[q + 1]) <= tie_rtol * max(ps[q], 1e-300) for q in range(K - 1)
                        if ps[q] > 0]
                if min(abs(ua[j] - c) for c in cdf) <= cdf_tol:
                    out["boundary"] += 1
                elif any(gaps):
                    out["near_tie"] += 1
                else:
                    out["hard"] += 1
                    if out["first_hard"] is None:
                        out["first_hard"] = {"it": it, "cell": j, "oracle": lab, "recorded": int(rec_lab[j]),
                                             "u": float(ua[j]), "prob": prob[j].tolist(),
                                             "rec_prob": rec_prob[j].tolist()}
        z = rec_lab - 1
        # (c) Config flips
        rec_cfg = np.asarray(rec["Config_all"][it - 1]).reshape(K, N).T
        Zm = np.zeros((M, K))
        Zm[np.arange(M), z] = 1.0
        SA, SB = A1 @ Zm, B1 @ Zm
        if relax:
            if it > (0.1 * min_iter + 5) and relax_rate_fixed is None:
                relax_rate = float(np.asarray(rec["relax_rate_all"]).reshape(-1)[it - 1])
            oddlog = _config_prior_oddlog(Config, relax_rate, keep_base_clone)
            odd = SA * (l1 - l0) + SB * (l1c - l0c) + oddlog
            odd = np.clip(odd, -50, 50)
            p = np.exp(odd) / (np.exp(odd) + 1)
            uc = u_config[it - 1] if u_config is not None else \
                philox.uniforms(seed, chain, it, philox.STREAM_CONFIG, np.arange(N * K))
            new = r_rbinom1(p.T.ravel(), uc).reshape(K, N).T
            bad = new != rec_cfg
            if bad.any():
                ucm = np.asarray(uc).reshape(K, N).T
                thr = np.where(p > 0.5, p, 1 - np.minimum(p, 1 - p))
                near = np.abs(ucm - thr) <= cdf_tol
                out["config_boundary"] += int((bad & near).sum())
                out["config_hard"] += int((bad & ~near).sum())
        else:
            out["config_hard"] += int((rec_cfg != Config).sum())
        # (d) logLik trace with the OLD theta and the RECORDED new Config
        if dense_loglik:
            ll = get_logLik(A1, B1, rec_cfg, rec_lab, theta0, np.broadcast_to(th1v.reshape(-1, 1), A1.shape))
        else:
            ll = get_logLik_tables(A1, B1, rec_cfg, rec_lab, theta0, th1v)
        rl = float(np.asarray(rec["logLik"]).reshape(-1)[it - 1])
        out["max_dloglik_rel"] = max(out["max_dloglik_rel"], abs(ll - rl) / max(abs(ll), 1e-300))
        Config_prev = rec_cfg
    return out


class RecordingDraws(NumpyDraws):
    """NumpyDraws that also records every uniform it hands out, so a chain can be replayed elsewhere
    (what tools/record_reference.R does for a genuine R run)."""

    def __init__(self, seed, T, M, NK):
        super().__init__(seed)
        self.ua = np.full((T, M), 0.5)
        self.uc = np.full((T, NK), 0.5)

    def u_assign(self, it, M):
        u = self.rng.random(M)
        self.ua[it - 1] = u
        return u

    def u_config(self, it, NK):
        u = self.rng.random(NK)
        self.uc[it - 1] = u
        return u
