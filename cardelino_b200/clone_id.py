"""Host-side mirror of the reference's R interface for the clone-assignment path.

Same names, argument meaning, defaults, error messages and return-list fields as
R/clone_id.R:78-81 (clone_id), :351-356 (clone_id_Gibbs), :240-242 (clone_id_EM); the sampler body is
a call through the C ABI (include/cardelino_b200.h) into the sm_100a kernels.  R itself binds the same
ABI through r/cardelino_b200_shim.c + r/clone_id_b200.R; this module exists because the
build image has no R, and so that the parity tests read like tests/testthat/test-clone_id.R.

Matrices are variants x cells like the reference; NA is NaN; pandas DataFrames carry dimnames
(index = variant ids, columns = cell ids) and are matched by variant id like R/clone_id.R:108-118.
Plain arrays have no dimnames and are taken as already aligned.
"""
import itertools
import math
import warnings

import numpy as np

from . import _lib

try:  # optional conveniences for dimnames / sparse input
    import pandas as _pd
except Exception:  # pragma: no cover
    _pd = None
try:
    import scipy.sparse as _sp
except Exception:  # pragma: no cover
    _sp = None

_DENSE_LIMIT = 2 * 10 ** 8     # N*M above which dense N x M outputs (theta1, prob_variant) are skipped

_default_handle = None


def _handle(device=0):
    global _default_handle
    if _default_handle is None:
        _default_handle = _lib.Handle(device)
    return _default_handle


class SparseCounts:
    """Compact host form of (A, D): one CSC pattern over cells (D > 0 everywhere), uint16 counts --
    what load_cellSNP_vcf's two dgCMatrix objects (R/read_data.R:293-294) reduce to.  Pass it as `A`
    (with D=None).  cell_offset / M_total mark a cell shard of a larger donor."""

    def __init__(self, N, M, indptr, indices, a, d, cell_offset=0, M_total=None):
        self.shape = (int(N), int(M))
        self.indptr, self.indices, self.a, self.d = indptr, indices, a, d
        self.cell_offset, self.M_total = int(cell_offset), int(M if M_total is None else M_total)

    @classmethod
    def from_sparse(cls, A, D):
        """Merge two SciPy sparse matrices (variants x cells) -- the AD and DP matrices cellSNP / vcf loading
        yields (R/read_data.R:293-294) -- into the one-pattern uint16 form, without densifying (replaces
        as.matrix() at R/clone_id.R:121-122).  D == 0 is missing; an A entry there is dropped and an A entry absent
        where D > 0 is 0 (R/clone_id.R:380-386)."""
        if _sp is None:
            raise ImportError("scipy is needed for sparse inputs")
        D = _sp.csc_matrix(D, dtype=np.float64, copy=True)
        A = _sp.csc_matrix(A, dtype=np.float64, copy=True)
        if A.shape != D.shape:
            raise ValueError("A and D must have the same size")
        for X in (A, D):
            X.sum_duplicates()
            X.eliminate_zeros()
            X.sort_indices()
        if np.any(A.data % 1 != 0) or np.any(D.data % 1 != 0):
            raise ValueError("A and/or D contain non-integer values.")
        if A.nnz and A.data.min() < 0 or D.nnz and D.data.min() < 0:
            raise ValueError("A and/or D contain negative values.")
        if (A.nnz and A.data.max() > 65535) or (D.nnz and D.data.max() > 65535):
            raise ValueError("count above 65535 (uint16 storage)")
        N, M = D.shape
        col = lambda X: np.repeat(np.arange(M, dtype=np.int64), np.diff(X.indptr))
        key_d = col(D) * N + D.indices
        key_a = col(A) * N + A.indices
        pos = np.searchsorted(key_d, key_a)
        pos[pos == key_d.size] = 0
        hit = key_d[pos] == key_a if key_d.size else np.zeros(key_a.size, dtype=bool)
        a = np.zeros(D.nnz, dtype=np.uint16)
        a[pos[hit]] = A.data[hit].astype(np.uint16)
        d = D.data.astype(np.uint16)
        if np.any(a > d):
            raise ValueError("A exceeds D at a covered entry")
        return cls(N, M, D.indptr.astype(np.int64), D.indices.astype(np.int32), a, d)

    @classmethod
    def from_mtx(cls, ad_path, dp_path):
        """The same from two MatrixMarket files (e.g. the reference's inst/extdata/passed_ad.mtx / passed_dp.mtx)."""
        from scipy.io import mmread

        def read(path):
            try:
                return mmread(path, spmatrix=False)      # SciPy >= 1.15: say which sparse flavour explicitly
            except TypeError:
                return mmread(path)
        return cls.from_sparse(read(ad_path), read(dp_path))

    def to_dense(self):
        """(A, D) as float64 arrays with NaN where D is missing -- the form sim_read_count emits."""
        N, M = self.shape
        A = np.full((N, M), np.nan)
        D = np.full((N, M), np.nan)
        cols = np.repeat(np.arange(M), np.diff(np.asarray(self.indptr)))
        A[self.indices, cols] = self.a
        D[self.indices, cols] = self.d
        return A, D


def _names(x):
    if _pd is not None and isinstance(x, _pd.DataFrame):
        return list(x.index), list(x.columns)
    return None, None


def _values(x):
    if _pd is not None and isinstance(x, _pd.DataFrame):
        return x.to_numpy(dtype=np.float64)
    if x is None or isinstance(x, SparseCounts) or (_sp is not None and _sp.issparse(x)):
        return x
    return np.asarray(x, dtype=np.float64)


def _load(h, A, D):
    """Replaces as.matrix() + NA normalisation (R/clone_id.R:121-122,380-391)."""
    if isinstance(A, SparseCounts):
        h.load_csc_u16(A.shape[0], A.shape[1], A.indptr, A.indices, A.a, A.d)
        if A.cell_offset or A.M_total != A.shape[1]:
            h.set_shard(A.cell_offset, A.M_total)
    elif _sp is not None and _sp.issparse(D):
        Dc = _sp.csc_matrix(D)
        Ac = _sp.csc_matrix(A)
        Dc.sort_indices()
        Ac.sort_indices()
        if Ac.shape != Dc.shape:
            raise ValueError("A and D must have the same size")
        h.load_csc(Dc.shape[0], Dc.shape[1], Dc.indptr, Dc.indices, Dc.data, Ac.indptr, Ac.indices, Ac.data)
    else:
        if _sp is not None and _sp.issparse(A):
            A = np.asarray(A.todense(), dtype=np.float64)
        h.load_dense(A, D)


def colMatch(A, B, force=False):
    """R/assessment.R:87-121 -- column indices (0-based) of B matched to A by minimum mean absolute
    difference.  force=TRUE enumerates all permutations in the reference (O(K!)); the same optimum is
    found here by enumeration for K <= 8 and by an exact assignment solver above that."""
    A = np.asarray(A, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    if A.shape[0] != B.shape[0]:
        raise ValueError("A and B have different rows.")
    if A.shape[1] > B.shape[1]:
        raise ValueError("A must have equal or smaller columns than B.")
    ka, kb = A.shape[1], B.shape[1]
    # ka x kb column distances, one column of A at a time (an M x ka x kb temporary is 2 GB at 10^6 cells x 16)
    cost = np.stack([np.abs(A[:, i:i + 1] - B).mean(axis=0) for i in range(ka)]) if ka else np.zeros((0, kb))
    if force:
        if kb <= 8:
            best, best_idx = None, None
            for perm in itertools.permutations(range(kb)):
                idx = perm[:ka]
                v = sum(cost[i, idx[i]] for i in range(ka))
                if best is None or v < best:
                    best, best_idx = v, idx
            return np.array(best_idx)
        from scipy.optimize import linear_sum_assignment
        r, c = linear_sum_assignment(cost)
        out = np.zeros(ka, dtype=int)
        out[r] = c
        return out
    tmp = cost.copy()
    out = -np.ones(ka, dtype=int)
    big = cost.max() + 1
    while (out == -1).any():
        flat = np.argmin(tmp.T.ravel())
        c, r = divmod(flat, ka)
        out[r] = c
        tmp[r, :] = big
        tmp[:, c] = big
    return out


def devianceIC(logLik_all, logLik_post, verbose=True):
    """R/clone_id.R:765-798."""
    logLik_all = np.asarray(logLik_all, dtype=np.float64)
    m = float(np.mean(logLik_all))
    v = float(np.var(logLik_all, ddof=1)) if logLik_all.size > 1 else float("nan")
    p_D_S = -2 * m - (-2 * logLik_post)
    DIC_S = -2 * logLik_post + 2 * p_D_S
    DIC_G = -2 * logLik_post + 2 * (2 * v)
    if verbose:
        print("DIC: %s D_mean: %s D_post: %s logLik_var: %s " %
              (round(DIC_G, 2), round(-2 * m, 2), round(-2 * logLik_post, 2), round(v, 2)))
    return {"DIC": DIC_G, "logLik_var": v, "D_mean": -2 * m, "D_post": -2 * logLik_post,
            "DIC_Gelman": DIC_G, "DIC_Spiegelhalter": DIC_S}


def Geweke_Z(X, first=0.1, last=0.5):
    """R/clone_id.R:688-715 on a host matrix (small diagnostics; the sampler's stop rule runs on device)."""
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 1:
        X = X.reshape(-1, 1)
    n = X.shape[0]
    Aw = X[: int(math.floor(first * n))]
    Bw = X[int(math.ceil(last * n)) - 1:]
    with np.errstate(invalid="ignore", divide="ignore"):
        mA, mB = Aw.mean(axis=0), Bw.mean(axis=0)
        vA = ((Aw - mA) ** 2).sum(axis=0) / (Aw.shape[0] - 1)
        vB = ((Bw - mB) ** 2).sum(axis=0) / (Aw.shape[0] - 1)
        return (mA - mB) / np.sqrt(vA + vB + 1e-50)


def _check_shapes(A, D, Config, Psi, gibbs):
    if D is None and isinstance(A, SparseCounts):
        D = A
    if A.shape != D.shape or A.shape[0] != Config.shape[0] or Config.shape[1] != len(Psi):
        raise ValueError("A and D must have the same size;\n" + (" " if gibbs else "") +
                         "A and Config must have the same number of variants;\n"
                         "Config and Psi must have the same number of clones")


def get_logLik(A, D, Config, Assign, theta0, theta1, handle=None):
    """R/clone_id.R:734-752 on the device.  `theta1` is a scalar or one value per variant."""
    h = handle or _handle()
    _load(h, _values(A), _values(D))
    return h.get_loglik(Config, Assign, theta0, theta1)


def clone_id_EM(A, D, Config, Psi=None, min_iter=10, max_iter=1000, logLik_threshold=1e-5, verbose=True,
                theta_init=None, seed=1, handle=None):
    """R/clone_id.R:240-326.  Returns {"theta", "prob", "logLik"} (+ "n_iter").  `theta_init` injects the
    two runif() starts of :278 (default: Philox stream of `seed`)."""
    Av, Dv = _values(A), _values(D)
    Config = np.asarray(_values(Config), dtype=np.float64)
    if Psi is None:
        Psi = np.full(Config.shape[1], 1.0 / Config.shape[1])
    _check_shapes(Av, Dv, Config, Psi, gibbs=False)
    h = handle or _handle()
    _load(h, Av, Dv)
    out = h.em_run(Config, Psi, min_iter, max_iter, logLik_threshold, theta_init, seed)
    if verbose:
        print("Total iterations: %d" % out["n_iter"])
    return out


def clone_id_Gibbs(A, D, Config, Psi=None, relax_Config=True, relax_rate_fixed=None, relax_rate_prior=(1, 9),
                   keep_base_clone=True, prior0=(0.2, 99.8), prior1=(0.45, 0.55), min_iter=5000,
                   max_iter=20000, buin_frac=0.5, wise="variant", relabel=False, verbose=True,
                   seed=1, n_chain=1, replay=None, keep_assign_all=False, keep_prob_all=-1,
                   handle=None, light=None, incremental_stats=None, chain_offset=0, _loaded=False,
                   _all_chains=False, progress=None, **lib_kw):
    """R/clone_id.R:351-675.  Returns the reference's list (:662-673) as a dict.  Extra arguments: `seed`
    (Philox key; the reference uses R's global RNG), `n_chain` (chains run inside the library), `replay`
    (dict of injected draws, see cdl_replay), `light` (leave out the fields that are dense N x M or
    max_iter x N*K doubles -- theta1, prob_variant, element, Config_all, prob_all; default: only when they would
    not fit a workstation's memory, i.e. N*M > 2e8 or the Config trace above 2 GB), `incremental_stats` (None: the
    library default, on wherever it applies; see cdl_gibbs_params), `progress(chain, iteration, fraction)` (called
    every 100 sweeps; a true return value stops the run -- `verbose` prints the reference's "x% converged." lines
    through it)."""
    Av, Dv = _values(A), _values(D)
    Config = np.asarray(_values(Config), dtype=np.float64)
    if Psi is None:
        Psi = np.full(Config.shape[1], 1.0 / Config.shape[1])
    _check_shapes(Av, Dv, Config, Psi, gibbs=True)
    if wise not in ("element", "variant", "global"):
        raise ValueError("Input wise mode: %s, while only supporting: element, variant, global" % wise)
    relax = relax_Config is not None and bool(relax_Config)        # R: !is.null(relax_Config) && relax_Config != FALSE
    if relax and relax_rate_fixed is not None and (relax_rate_fixed > 1 or relax_rate_fixed < 0):
        raise ValueError("relax_rate_fixed needs to be NULL or in [0, 1].")
    if relax and relax_rate_fixed is None and relax_rate_prior is None:
        raise ValueError("Require value for either relax_Config or relax_prior.")
    prior1 = np.asarray(prior1, dtype=np.float64)
    prior1_matrix = None
    if prior1.ndim == 2:
        if prior1.shape[1] != 2 or prior1.shape[0] < 1:
            raise ValueError("prior1 need to be a matrix of n_element x 2")
        prior1_matrix = prior1
        prior1 = prior1[0]
    elif not (prior1.ndim == 1 and prior1.size == 2):
        raise ValueError("prior1 need to be a matrix of n_element x 2")
    h = handle or _handle()
    if not _loaded:
        _load(h, Av, Dv)
    N, M = Av.shape
    K = Config.shape[1]
    if prior1_matrix is not None:
        # one row per theta1 element (R/clone_id.R:418-424): 1, N or the number of covered entries
        n_el = {"global": 1, "variant": N}.get(wise) or h.data_info()[2]
        if prior1_matrix.shape[0] != n_el:
            raise ValueError("prior1 need to be a matrix of n_element x 2 (n_element = %d for wise = %r, got %d rows)"
                             % (n_el, wise, prior1_matrix.shape[0]))
    if light is None:
        light = (not relabel) and (N * M > _DENSE_LIMIT or float(max_iter) * N * K * 8.0 > 2.0 ** 31)
    h.gibbs_run(Config, Psi, replay=replay, relax_Config=int(relax), relax_rate_fixed=relax_rate_fixed,
                relax_rate_prior=relax_rate_prior if relax_rate_prior is not None else (1, 9),
                keep_base_clone=int(bool(keep_base_clone)), prior0=prior0, prior1=prior1,
                prior1_matrix=prior1_matrix, min_iter=int(min_iter), max_iter=int(max_iter),
                buin_frac=float(buin_frac), wise=wise, n_chain=int(n_chain), seed=int(seed),
                keep_assign_all=int(bool(keep_assign_all)),
                keep_prob_all=1 if relabel else int(keep_prob_all), verbose=0, relabel=int(bool(relabel)),
                incremental_stats=-1 if incremental_stats is None else int(bool(incremental_stats)),
                chain_offset=int(chain_offset), progress=_progress_printer(verbose, progress), **lib_kw)
    outs = [_collect(h, c, Av, Config, wise, relabel, verbose, keep_assign_all, light) for c in range(n_chain)]
    return outs if _all_chains else outs[0]


def _progress_printer(verbose, user_cb):
    """The reference cat()s "x% converged." at every convergence check (R/clone_id.R:583-588); the library reports the
    fraction through its progress callback and prints nothing itself."""
    if not verbose and user_cb is None:
        return None

    def cb(chain, it, frac):
        if verbose and frac == frac:
            print("%g%% converged." % (round(frac, 3) * 100))
        return bool(user_cb(chain, it, frac)) if user_cb is not None else False
    return cb


def _collect(h, chain, Av, Config, wise, relabel, verbose, keep_assign_all, light=False):
    N, M = Av.shape
    if light:   # large donors: only the fields that stay small (no dense N x M, no big traces)
        info = h.gibbs_info(chain)
        pc = info.percent_converged
        if pc <= 95 or pc != pc:
            warnings.warn("Only %s%% of parameters converged. Consider increasing `max_iter`." % pc)
        elif verbose:
            print("Converged in %d iterations." % info.n_iter)
        theta0, theta1_el = h.fetch("theta", chain)
        theta0_all, theta1_all = h.fetch("theta_traces", chain)
        relax_rate, relax_rate_all = h.fetch("relax_rate", chain)
        return {"theta0": theta0, "theta1_elements": theta1_el, "theta0_all": theta0_all.reshape(-1, 1),
                "theta1_all": theta1_all, "logLik": h.fetch("logLik", chain), "prob": h.fetch("prob", chain),
                "relax_rate": relax_rate, "relax_rate_all": relax_rate_all,
                "Config_prob": h.fetch("Config_prob", chain), "n_iter": info.n_iter, "n_buin": info.n_buin,
                "percent_converged": info.percent_converged, "logLik_post": info.logLik_post,
                "W_log": info.W_log, "exact_trace": bool(info.exact_trace), "stop_rule": int(info.stop_rule),
                "DIC": dict(zip(("DIC", "logLik_var", "D_mean", "D_post", "DIC_Gelman", "DIC_Spiegelhalter"),
                                list(info.dic)))}
    K = Config.shape[1]
    info = h.gibbs_info(chain)
    it, n_buin = info.n_iter, info.n_buin
    pc = info.percent_converged
    if pc <= 95 or pc != pc:
        warnings.warn("Only %s%% of parameters converged. Consider increasing `max_iter`." % pc)
    elif verbose:
        print("Converged in %d iterations." % it)
    theta0, theta1_el = h.fetch("theta", chain)
    theta0_all, theta1_all = h.fetch("theta_traces", chain)
    relax_rate, relax_rate_all = h.fetch("relax_rate", chain)
    out = {
        "theta0": theta0, "theta0_all": theta0_all.reshape(-1, 1), "theta1_all": theta1_all,
        "logLik": h.fetch("logLik", chain), "prob": h.fetch("prob", chain),
        "relax_rate": relax_rate, "Config_prob": h.fetch("Config_prob", chain),
        "Config_all": h.fetch("Config_all", chain), "relax_rate_all": relax_rate_all,
        "n_iter": it, "n_buin": n_buin, "percent_converged": pc, "logLik_post": info.logLik_post,
        "W_log": info.W_log, "exact_trace": bool(info.exact_trace), "stop_rule": int(info.stop_rule),
    }
    if info.exact_trace:
        out["prob_all"] = h.fetch("prob_all", chain)
    if keep_assign_all:
        out["assign_all"] = h.fetch("assign_all", chain)
    if N * M <= _DENSE_LIMIT:
        if wise == "element":
            el = h.fetch("element")                            # which(!is.na(D)), column-major (:413-415)
            th = np.full(N * M, np.nan)                        # theta1 stays NA off the covered entries (:463,653)
            th[el] = theta1_el
            out["element"] = el
            out["theta1"] = th.reshape(M, N).T.copy()
        else:
            out["element"] = np.arange(N * M)                  # idx_mat for global / variant (:408,411)
            out["theta1"] = (np.full((N, M), theta1_el[0]) if wise == "global"
                             else np.repeat(theta1_el.reshape(N, 1), M, axis=1))
        out["prob_variant"] = h.fetch("prob_variant", chain)
    else:
        out["theta1_elements"] = theta1_el
    logLik_post = info.logLik_post       # relabel (R/clone_id.R:624-644) ran on the device before the summaries
    out["DIC"] = devianceIC(out["logLik"][n_buin - 1:it], logLik_post, verbose=verbose)
    return out


def merge_chains(outs):
    """Chain merge of R/clone_id.R:165-181: match every chain's clone columns to chain 1 with colMatch(force=TRUE),
    then average prob, relax_rate and Config_prob; every other field is chain 1's.  `outs` are per-chain results
    (dicts with at least prob, relax_rate, Config_prob) -- from one GPU or gathered from several."""
    out = dict(outs[0])
    out["n_chain"] = 1
    n_chain = len(outs)
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


def clone_id(A, D, Config=None, n_clone=None, Psi=None, relax_Config=True, relax_rate_fixed=None,
             inference="sampling", n_chain=1, n_proc=1, verbose=True, handle=None, **kwargs):
    """R/clone_id.R:78-186.  `n_proc` is accepted for signature compatibility; chains run inside the
    library on the GPU (a forked doMC worker cannot share a CUDA context)."""
    if inference not in ("sampling", "EM"):
        raise ValueError("'arg' should be one of 'sampling', 'EM'")
    a_rows, a_cols = _names(A)
    d_rows, d_cols = _names(D)
    if a_rows is not None and d_rows is not None:
        if a_rows != d_rows:
            raise ValueError("Rownames for A and D are not identical.")
        if a_cols != d_cols:
            raise ValueError("Colnames for A and D are not identical.")
    if Config is None and n_clone is None:
        raise ValueError("Config and n_clone can't be NULL together.")
    Av, Dv = _values(A), _values(D)
    if isinstance(Av, SparseCounts) and Dv is None:
        Dv = Av                 # compact counts carry both matrices (already validated integers)
    for name, X in (("A", Av), ("D", Dv)):
        if isinstance(X, SparseCounts):
            continue
        vals = X.data if (_sp is not None and _sp.issparse(X)) else X
        with np.errstate(invalid="ignore"):
            if np.any((vals % 1 != 0) & ~np.isnan(vals)):
                raise ValueError("A and/or D contain non-integer values.")
    cell_names = d_cols
    if Config is None:
        print("Config is NULL: de-novo mode is in use.")
        Cv = np.zeros((Dv.shape[0], int(n_clone)))
        c_rows, c_cols = d_rows, ["Clone%d" % (k + 1) for k in range(int(n_clone))]
        relax_Config = True
        relax_rate_fixed = 0.5
    else:
        c_rows, c_cols = _names(Config)
        Cv = np.asarray(_values(Config), dtype=np.float64)
    if c_rows is not None and d_rows is not None:
        d_set = set(d_rows)
        common = [v for v in c_rows if v in d_set]                 # intersect(rownames(Config), rownames(D))
        if not common:
            raise ValueError("No matches in variant names between Config and D arguments.")
        di = {v: n for n, v in enumerate(d_rows)}
        ci = {v: n for n, v in enumerate(c_rows)}
        rows_d = [di[v] for v in common]
        Av, Dv = Av[rows_d, :], Dv[rows_d, :]
        Cv = Cv[[ci[v] for v in common], :]
        n_used = len(common)
    else:
        n_used = Dv.shape[0]
    if verbose:
        print("%d variants used for cell assignment." % n_used)
    if inference == "EM":
        out = clone_id_EM(Av, Dv, Cv, verbose=verbose, handle=handle, **kwargs)
    else:
        outs = clone_id_Gibbs(Av, Dv, Cv, Psi=Psi, relax_Config=relax_Config, relax_rate_fixed=relax_rate_fixed,
                              verbose=verbose, n_chain=n_chain, handle=handle, _all_chains=True, **kwargs)
        out = merge_chains(outs)
    out["cell_names"] = cell_names
    out["clone_names"] = c_cols
    return out


def assign_cells_to_clones(prob_mat, threshold=0.5, cell_names=None, clone_names=None):
    """R/clone_id.R:208-222 (post-processing on the returned prob; unchanged semantics)."""
    prob_mat = np.asarray(prob_mat)
    M, K = prob_mat.shape
    cell_names = cell_names if cell_names is not None else ["cell%d" % (j + 1) for j in range(M)]
    clone_names = clone_names if clone_names is not None else ["clone%d" % (k + 1) for k in range(K)]
    pmax = prob_mat.max(axis=1)
    arg = prob_mat.argmax(axis=1)
    clone = [clone_names[arg[j]] if pmax[j] > threshold else "unassigned" for j in range(M)]
    return {"cell": list(cell_names), "clone": clone, "prob_max": pmax}
