"""ctypes binding of libcardelino_b200.so (include/cardelino_b200.h).

There is no CPU fallback: if the CUDA library is missing or no B200-class device is visible, every
entry point raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# CARDELINO_B200_LIB selects another build of the same library (kernel A/B experiments); the default is the in-tree one
LIB_PATH = os.environ.get("CARDELINO_B200_LIB") or os.path.join(_HERE, "libcardelino_b200.so")

EXPORTS = [
    "cdl_version", "cdl_last_error", "cdl_create", "cdl_destroy", "cdl_load_dense", "cdl_load_csc",
    "cdl_load_csc_u16", "cdl_synth_generate", "cdl_synth_labels", "cdl_export_csc_u16", "cdl_data_info", "cdl_set_shard",
    "cdl_comm_unique_id", "cdl_comm_init", "cdl_gibbs_defaults", "cdl_gibbs_run", "cdl_gibbs_info_get",
    "cdl_fetch_prob", "cdl_fetch_config_prob", "cdl_fetch_theta", "cdl_fetch_theta_traces",
    "cdl_fetch_loglik_trace", "cdl_fetch_relax_rate", "cdl_fetch_prob_all", "cdl_fetch_config_all",
    "cdl_fetch_assign_all", "cdl_fetch_prob_variant", "cdl_fetch_element_index", "cdl_get_loglik",
    "cdl_em_run", "cdl_bench_sweeps", "cdl_set_cells_layout", "cdl_set_exchange_mode", "cdl_exchange_is_peer", "cdl_incremental_full_passes",
    "cdl_debug_capture", "cdl_debug_fetch",
]

WISE = {"global": 0, "variant": 1, "element": 2}
CELLS_LAYOUT = {"auto": 0, "rows": 1, "slices": 2, "small": 3}


class CdlError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


# int (*cdl_progress_fn)(void* user, int32_t chain, int32_t iteration, double fraction)
PROGRESS_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int32, C.c_int32, C.c_double)


class GibbsParams(C.Structure):
    _fields_ = [
        ("relax_Config", C.c_int32), ("relax_rate_fixed_set", C.c_int32), ("relax_rate_fixed", C.c_double),
        ("relax_rate_prior", C.c_double * 2), ("keep_base_clone", C.c_int32), ("prior0", C.c_double * 2),
        ("prior1", C.c_double * 2), ("prior1_matrix", C.POINTER(C.c_double)), ("min_iter", C.c_int32),
        ("max_iter", C.c_int32), ("buin_frac", C.c_double), ("wise", C.c_int32), ("n_chain", C.c_int32),
        ("seed", C.c_uint64), ("keep_assign_all", C.c_int32), ("keep_prob_all", C.c_int32),
        ("trace_budget_bytes", C.c_int64), ("check_every", C.c_int32), ("verbose", C.c_int32),
        ("acc_fp32", C.c_int32), ("chain_offset", C.c_int32), ("incremental_stats", C.c_int32),
        ("relabel", C.c_int32), ("prior1_matrix_rows", C.c_int64), ("progress", PROGRESS_FN),
        ("progress_user", C.c_void_p), ("snapshot_fp32", C.c_int32),
    ]


class Replay(C.Structure):
    _fields_ = [
        ("n_rows", C.c_int32), ("u_assign", C.POINTER(C.c_double)), ("u_config", C.POINTER(C.c_double)),
        ("theta0", C.POINTER(C.c_double)), ("theta1", C.POINTER(C.c_double)),
        ("relax_rate", C.POINTER(C.c_double)),
    ]


class GibbsInfo(C.Structure):
    _fields_ = [
        ("n_iter", C.c_int32), ("n_buin", C.c_int32), ("n_element", C.c_int32), ("exact_trace", C.c_int32),
        ("percent_converged", C.c_double), ("logLik_post", C.c_double), ("W_log", C.c_double),
        ("dic", C.c_double * 6), ("stop_rule", C.c_int32),
    ]


_lib = None


def load_library():
    """dlopen the in-tree CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CdlError(2, "libcardelino_b200.so is not built (run `python -m cardelino_b200.build`); "
                          "there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    lib.cdl_last_error.restype = C.c_char_p
    lib.cdl_last_error.argtypes = [C.c_void_p]
    lib.cdl_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.cdl_destroy.argtypes = [C.c_void_p]
    lib.cdl_gibbs_defaults.restype = None
    _lib = lib
    return lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a, order="F"):
    return np.require(np.asarray(a, dtype=np.float64), requirements=["A", "F" if order == "F" else "C"])


class Handle:
    """Opaque device-side state (donor layouts, chains) behind the C ABI."""

    def __init__(self, device=0):
        self.lib = load_library()
        self.h = C.c_void_p()
        rc = self.lib.cdl_create(int(device), C.byref(self.h))
        if rc != 0:
            raise CdlError(rc, self.lib.cdl_last_error(None).decode())
        self.device = device

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.cdl_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise CdlError(rc, self.lib.cdl_last_error(self.h).decode())

    # ---- data ----
    def load_dense(self, A, D):
        A = _f64(A)
        D = _f64(D)
        if A.shape != D.shape:
            raise CdlError(1, "A and D must have the same size")
        self._check(self.lib.cdl_load_dense(self.h, C.c_int64(A.shape[0]), C.c_int64(A.shape[1]), _dp(A), _dp(D)))

    def load_csc(self, N, M, d_p, d_i, d_x, a_p, a_i, a_x):
        arrs = [np.ascontiguousarray(d_p, np.int64), np.ascontiguousarray(d_i, np.int32),
                np.ascontiguousarray(d_x, np.float64), np.ascontiguousarray(a_p, np.int64),
                np.ascontiguousarray(a_i, np.int32), np.ascontiguousarray(a_x, np.float64)]
        self._check(self.lib.cdl_load_csc(self.h, C.c_int64(N), C.c_int64(M), *[C.c_void_p(x.ctypes.data) for x in arrs]))

    def load_csc_u16(self, N, M, p, i, a, d):
        arrs = [np.ascontiguousarray(p, np.int64), np.ascontiguousarray(i, np.int32),
                np.ascontiguousarray(a, np.uint16), np.ascontiguousarray(d, np.uint16)]
        self._check(self.lib.cdl_load_csc_u16(self.h, C.c_int64(N), C.c_int64(M),
                                              *[C.c_void_p(x.ctypes.data) for x in arrs]))

    def set_cells_layout(self, layout):
        """'auto' | 'rows' | 'slices': which cell kernel the Gibbs sweep uses (include/cardelino_b200.h)."""
        self._check(self.lib.cdl_set_cells_layout(self.h, C.c_int32(CELLS_LAYOUT[layout])))

    def set_exchange_mode(self, mode):
        """'auto' | 'nccl' | 'peer': how the ranks of a cell-sharded run sum the statistic table per sweep."""
        self._check(self.lib.cdl_set_exchange_mode(self.h, C.c_int32({"auto": 0, "nccl": 1, "peer": 2}[mode])))

    def exchange_is_peer(self):
        return bool(self.lib.cdl_exchange_is_peer(self.h))

    def set_shard(self, cell_offset, M_total):
        self._check(self.lib.cdl_set_shard(self.h, C.c_int64(cell_offset), C.c_int64(M_total)))

    def synth_generate(self, N, M_total, K, Config, coverage, seed, cell_offset=0, M_local=None):
        Config = _f64(Config)
        M_local = M_total if M_local is None else M_local
        self._check(self.lib.cdl_synth_generate(self.h, C.c_int64(N), C.c_int64(M_total), C.c_int64(cell_offset),
                                                C.c_int64(M_local), C.c_int32(K), _dp(Config),
                                                C.c_double(coverage), C.c_uint64(seed)))

    def synth_labels(self):
        N, M, nnz, _ = self.data_info()
        out = np.zeros(M, np.int32)
        self._check(self.lib.cdl_synth_labels(self.h, C.c_void_p(out.ctypes.data)))
        return out

    def data_info(self):
        N, M, nnz, W = C.c_int64(), C.c_int64(), C.c_int64(), C.c_double()
        self._check(self.lib.cdl_data_info(self.h, C.byref(N), C.byref(M), C.byref(nnz), C.byref(W)))
        return N.value, M.value, nnz.value, W.value

    def export_csc_u16(self):
        N, M, nnz, _ = self.data_info()
        p = np.zeros(M + 1, np.int64)
        i = np.zeros(nnz, np.int32)
        a = np.zeros(nnz, np.uint16)
        d = np.zeros(nnz, np.uint16)
        self._check(self.lib.cdl_export_csc_u16(self.h, *[C.c_void_p(x.ctypes.data) for x in (p, i, a, d)]))
        return p, i, a, d

    # ---- comm ----
    def comm_unique_id(self, nccl_path):
        buf = np.zeros(128, np.uint8)
        self._check(self.lib.cdl_comm_unique_id(self.h, nccl_path.encode(), C.c_void_p(buf.ctypes.data)))
        return buf

    def comm_init(self, nccl_path, uid, rank, world):
        uid = np.ascontiguousarray(uid, np.uint8)
        self._check(self.lib.cdl_comm_init(self.h, nccl_path.encode(), C.c_void_p(uid.ctypes.data),
                                           C.c_int32(rank), C.c_int32(world)))

    # ---- Gibbs ----
    def gibbs_run(self, Config, Psi=None, replay=None, **kw):
        Config = _f64(Config)
        N, K = Config.shape
        p = GibbsParams()
        self.lib.cdl_gibbs_defaults(C.byref(p))
        keep = []
        pending = []
        for name, val in kw.items():
            if name in ("relax_rate_prior", "prior0", "prior1"):
                arr = getattr(p, name)
                arr[0], arr[1] = float(val[0]), float(val[1])
            elif name == "prior1_matrix":
                if val is not None:
                    m = _f64(val)
                    if m.ndim != 2 or m.shape[1] != 2:
                        raise CdlError(1, "prior1 must be a matrix with two columns")
                    keep.append(m)
                    p.prior1_matrix = _dp(m)
                    p.prior1_matrix_rows = m.shape[0]
            elif name == "progress":
                if val is not None:
                    # exceptions must not cross the C frames: remember the first one, ask the library to stop
                    # (CDL_ERR_INTERRUPTED) and re-raise after the call has returned
                    def _cb(_user, chain, it, frac, _f=val):
                        try:
                            return int(bool(_f(chain, it, frac)))
                        except BaseException as e:      # noqa: BLE001 (KeyboardInterrupt included)
                            if not pending:
                                pending.append(e)
                            return 1
                    fn = PROGRESS_FN(_cb)
                    keep.append(fn)
                    p.progress = fn
            elif name == "relax_rate_fixed":
                if val is None:
                    p.relax_rate_fixed_set = 0
                else:
                    p.relax_rate_fixed_set = 1
                    p.relax_rate_fixed = float(val)
            elif name == "wise":
                p.wise = WISE[val] if isinstance(val, str) else int(val)
            elif hasattr(p, name):
                setattr(p, name, type(getattr(p, name))(val))
            else:
                raise TypeError("unknown Gibbs parameter %r" % name)
        psi = None if Psi is None else _f64(Psi)
        rp = None
        if replay is not None:
            rp = Replay()
            rp.n_rows = int(replay["n_rows"])
            for f in ("u_assign", "u_config", "theta0", "theta1", "relax_rate"):
                v = replay.get(f)
                if v is not None:
                    v = np.ascontiguousarray(v, np.float64)
                    keep.append(v)
                    setattr(rp, f, _dp(v))
        rc = self.lib.cdl_gibbs_run(self.h, _dp(Config), C.c_int32(K), _dp(psi), C.byref(p),
                                    C.byref(rp) if rp is not None else None)
        if pending:
            raise pending[0]
        self._check(rc)
        self._N, self._K = N, K
        return p

    # ---- test hooks ----
    def debug_capture(self, enable=True):
        self._check(self.lib.cdl_debug_capture(self.h, C.c_int32(1 if enable else 0)))

    def debug_fetch(self):
        """(SA, SB) this rank's N x K statistic table of the last sweep (before any exchange), the integers added to
        the theta1 priors (S3[N], S4[N]) and the totals (S1, S2, S3, S4, diff1)."""
        N, K = self._N, self._K
        SA = np.zeros((N, K), np.uint64)
        SB = np.zeros((N, K), np.uint64)
        S3 = np.zeros(N, np.uint64)
        S4 = np.zeros(N, np.uint64)
        tot = np.zeros(5, np.uint64)
        self._check(self.lib.cdl_debug_fetch(self.h, *[C.c_void_p(x.ctypes.data) for x in (SA, SB, S3, S4, tot)]))
        return SA, SB, S3, S4, tot

    def incremental_full_passes(self, chain=0):
        out = C.c_uint32()
        self._check(self.lib.cdl_incremental_full_passes(self.h, C.c_int32(chain), C.byref(out)))
        return out.value

    def gibbs_info(self, chain=0):
        info = GibbsInfo()
        self._check(self.lib.cdl_gibbs_info_get(self.h, C.c_int32(chain), C.byref(info)))
        return info

    def fetch(self, what, chain=0):
        """Fields of the reference return list (R/clone_id.R:662-673) as numpy arrays in R's shapes."""
        N, M, nnz, _ = self.data_info()
        K = self._K
        info = self.gibbs_info(chain)
        T, n_el = info.n_iter, info.n_element
        ch = C.c_int32(chain)
        if what == "prob":
            out = np.zeros((M, K), order="F")
            self._check(self.lib.cdl_fetch_prob(self.h, ch, _dp(out)))
        elif what == "Config_prob":
            out = np.zeros((N, K), order="F")
            self._check(self.lib.cdl_fetch_config_prob(self.h, ch, _dp(out)))
        elif what == "theta":
            t0 = C.c_double()
            t1 = np.zeros(n_el)
            self._check(self.lib.cdl_fetch_theta(self.h, ch, C.byref(t0), _dp(t1)))
            out = (t0.value, t1)
        elif what == "theta_traces":
            t0 = np.zeros(T)
            t1 = np.zeros((T, n_el), order="F")
            self._check(self.lib.cdl_fetch_theta_traces(self.h, ch, _dp(t0), _dp(t1)))
            out = (t0, t1)
        elif what == "logLik":
            out = np.zeros(T)
            self._check(self.lib.cdl_fetch_loglik_trace(self.h, ch, _dp(out)))
        elif what == "relax_rate":
            m = C.c_double()
            al = np.zeros(T)
            self._check(self.lib.cdl_fetch_relax_rate(self.h, ch, C.byref(m), _dp(al)))
            out = (m.value, al)
        elif what == "prob_all":
            out = np.zeros((T, M * K))
            self._check(self.lib.cdl_fetch_prob_all(self.h, ch, _dp(out)))
        elif what == "Config_all":
            out = np.zeros((T, N * K))
            self._check(self.lib.cdl_fetch_config_all(self.h, ch, _dp(out)))
        elif what == "assign_all":
            out = np.zeros((T, M), np.int32)
            self._check(self.lib.cdl_fetch_assign_all(self.h, ch, C.c_void_p(out.ctypes.data)))
        elif what == "prob_variant":
            out = np.zeros((N, M), order="F")
            self._check(self.lib.cdl_fetch_prob_variant(self.h, ch, _dp(out)))
        elif what == "element":
            out = np.zeros(nnz, np.int64)
            self._check(self.lib.cdl_fetch_element_index(self.h, C.c_void_p(out.ctypes.data)))
        else:
            raise KeyError(what)
        return out

    # ---- deterministic pieces ----
    def get_loglik(self, Config, Assign, theta0, theta1):
        Config = _f64(Config)
        N, K = Config.shape
        Assign = np.asarray(Assign)
        th1 = np.atleast_1d(np.asarray(theta1, np.float64)).ravel().copy()
        out = C.c_double()
        if Assign.ndim == 1:
            a = np.ascontiguousarray(Assign, np.int32)
            rc = self.lib.cdl_get_loglik(self.h, _dp(Config), C.c_int32(K), C.c_void_p(a.ctypes.data), None,
                                         C.c_double(theta0), _dp(th1), C.c_int64(th1.size), C.byref(out))
        else:
            P = _f64(Assign)
            rc = self.lib.cdl_get_loglik(self.h, _dp(Config), C.c_int32(K), None, _dp(P), C.c_double(theta0),
                                         _dp(th1), C.c_int64(th1.size), C.byref(out))
        self._check(rc)
        return out.value

    def em_run(self, Config, Psi=None, min_iter=10, max_iter=1000, logLik_threshold=1e-5, theta_init=None, seed=1):
        Config = _f64(Config)
        N, K = Config.shape
        _, M, _, _ = self.data_info()
        psi = None if Psi is None else _f64(Psi)
        ti = None if theta_init is None else np.ascontiguousarray(theta_init, np.float64)
        theta = np.zeros(2)
        prob = np.zeros((M, K), order="F")
        ll = C.c_double()
        nit = C.c_int32()
        self._check(self.lib.cdl_em_run(self.h, _dp(Config), C.c_int32(K), _dp(psi), C.c_int32(min_iter),
                                        C.c_int32(max_iter), C.c_double(logLik_threshold), _dp(ti),
                                        C.c_uint64(seed), _dp(theta), _dp(prob), C.byref(ll), C.byref(nit)))
        return {"theta": theta, "prob": prob, "logLik": ll.value, "n_iter": nit.value}

    def bench_sweeps(self, warmup, n, breakdown=True):
        tot, tc, ts, tt = C.c_float(), C.c_float(), C.c_float(), C.c_float()
        launches = C.c_int64()
        self._check(self.lib.cdl_bench_sweeps(self.h, C.c_int32(warmup), C.c_int32(n), C.byref(tot),
                                              C.byref(tc) if breakdown else None,
                                              C.byref(ts) if breakdown else None,
                                              C.byref(tt) if breakdown else None, C.byref(launches)))
        return {"ms_total": tot.value, "ms_cells": tc.value, "ms_stats": ts.value, "ms_table": tt.value,
                "launches": launches.value}
