"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/gibbs_sparse.c (the sparse C / OpenMP restatement).

Built on demand with gcc into oracle/_build/ (git-ignored; it travels to the GPU box with the snapshot).  Used by
tests/ (pinned against the NumPy restatement, then used to check the CUDA path at BASELINE's full sizes) and by
bench.py's cpu_baseline / --impl reference legs.  Never imported by cardelino_b200/."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "gibbs_sparse.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libgibbs_sparse.so")

# quantiles (q + 0.5) / 256 of the positive depths of the packaged D_input (tools/rdata_to_npz.py); the same
# table as cardelino_b200/csrc/synth.cu
DEPTH_Q = np.array([v for v, n in [(1, 98), (2, 49), (3, 28), (4, 18), (5, 11), (6, 9), (7, 6), (8, 5), (9, 4), (10, 3), (11, 3), (12, 3), (13, 2), (14, 2), (15, 1), (16, 1), (17, 1), (18, 1), (19, 1), (21, 1), (22, 1), (24, 1), (27, 1), (30, 1), (34, 1), (38, 1), (49, 1), (61, 1), (107, 1)] for _ in range(n)], dtype=np.uint8)
assert DEPTH_Q.size == 256

_lib = None


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.run(["gcc", "-O3", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"], check=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_run.restype = C.c_double
        _lib.orc_synth.restype = C.c_int64
        _lib.orc_threads.restype = C.c_int
    return _lib


def _ptr(x):
    return C.c_void_p(x.ctypes.data) if x is not None else None


def threads():
    return int(lib().orc_threads())


def set_threads(n=None):
    """Use `n` OpenMP threads (default: every core this process may run on), whatever OMP_NUM_THREADS says."""
    if n is None:
        try:
            n = len(os.sched_getaffinity(0))
        except AttributeError:
            n = os.cpu_count() or 1
    lib().orc_set_threads(C.c_int(int(n)))
    return threads()


def config_masks(Config):
    Config = np.asarray(Config)
    m = np.zeros(Config.shape[0], np.uint32)
    for k in range(Config.shape[1]):
        m |= (Config[:, k] != 0).astype(np.uint32) << np.uint32(k)
    return m


def masks_to_config(mask, K):
    return ((mask[:, None] >> np.arange(K, dtype=np.uint32)[None, :]) & 1).astype(np.float64)


def sweep_cells(N, M, K, p, vi, a, d, mask, theta0, theta1, Psi, u_assign, want_prob=True):
    """R/clone_id.R:470-497 + the read sums SA / SB of the drawn labels.  Returns prob (M x K), z (0-based), SA, SB."""
    p = np.ascontiguousarray(p, np.int64); vi = np.ascontiguousarray(vi, np.int32)
    a = np.ascontiguousarray(a, np.uint16); d = np.ascontiguousarray(d, np.uint16)
    mask = np.ascontiguousarray(mask, np.uint32)
    th1 = np.ascontiguousarray(np.atleast_1d(theta1), np.float64)
    logPsi = np.ascontiguousarray(np.log(np.asarray(Psi, np.float64)))
    u = np.ascontiguousarray(u_assign, np.float64)
    prob = np.zeros((M, K)) if want_prob else None
    z = np.zeros(M, np.uint8)
    SA = np.zeros(N * K, np.uint64); SB = np.zeros(N * K, np.uint64)
    lib().orc_sweep_cells(C.c_int64(N), C.c_int64(M), C.c_int(K), _ptr(p), _ptr(vi), _ptr(a), _ptr(d), _ptr(mask),
                          C.c_double(theta0), _ptr(th1), C.c_int64(th1.size), _ptr(logPsi), _ptr(u), _ptr(prob), _ptr(z),
                          _ptr(SA), _ptr(SB))
    return prob, z, SA.reshape(N, K), SB.reshape(N, K)


def sweep_table(SA, SB, cfg0_mask, relax, keep_base, relax_rate, theta0, theta1, u_config, W_log):
    """R/clone_id.R:519-535,546-562,574-577 with the OLD theta.  Returns new masks, S1, S2, S3[N], S4[N], logLik, diff1."""
    N, K = SA.shape
    SA = np.ascontiguousarray(SA, np.uint64); SB = np.ascontiguousarray(SB, np.uint64)
    th1 = np.ascontiguousarray(np.atleast_1d(theta1), np.float64)
    uc = np.ascontiguousarray(u_config, np.float64)
    cfg0 = np.ascontiguousarray(cfg0_mask, np.uint32)
    out_mask = np.zeros(N, np.uint32)
    stats = np.zeros(2 + 2 * N)
    ll = C.c_double(); diff1 = C.c_int64()
    lib().orc_sweep_table(C.c_int64(N), C.c_int(K), _ptr(SA), _ptr(SB), _ptr(cfg0), C.c_int(int(relax)),
                          C.c_int(int(keep_base)), C.c_double(relax_rate), C.c_double(theta0), _ptr(th1),
                          C.c_int64(th1.size), _ptr(uc), C.c_double(W_log), _ptr(out_mask), _ptr(stats), C.byref(ll),
                          C.byref(diff1))
    return out_mask, stats[0], stats[1], stats[2:2 + N], stats[2 + N:], ll.value, diff1.value


def run(N, M, K, p, vi, a, d, cfg0_mask, T, seed=1):
    """T sweeps with the C oracle's own RNG; returns (seconds in the sweeps, final labels, final theta1)."""
    p = np.ascontiguousarray(p, np.int64); vi = np.ascontiguousarray(vi, np.int32)
    a = np.ascontiguousarray(a, np.uint16); d = np.ascontiguousarray(d, np.uint16)
    cfg0 = np.ascontiguousarray(cfg0_mask, np.uint32)
    z = np.zeros(M, np.uint8); th1 = np.zeros(N)
    sec = lib().orc_run(C.c_int64(N), C.c_int64(M), C.c_int(K), _ptr(p), _ptr(vi), _ptr(a), _ptr(d), _ptr(cfg0),
                        C.c_int(T), C.c_uint64(seed), _ptr(z), _ptr(th1))
    return float(sec), z, th1


def synth(N, M, K, cov, cfg_mask, seed=1):
    """sim_read_count-shaped donor generated on the host (same distributions as the device generator)."""
    cfg = np.ascontiguousarray(cfg_mask, np.uint32)
    p = np.zeros(M + 1, np.int64)
    L = lib()
    nnz = L.orc_synth(C.c_int64(N), C.c_int64(M), C.c_int(K), C.c_double(cov), _ptr(cfg), _ptr(DEPTH_Q), C.c_uint64(seed),
                      _ptr(p), None, None, None, None)
    vi = np.zeros(nnz, np.int32); a = np.zeros(nnz, np.uint16); d = np.zeros(nnz, np.uint16)
    labels = np.zeros(M, np.int32)
    L.orc_synth(C.c_int64(N), C.c_int64(M), C.c_int(K), C.c_double(cov), _ptr(cfg), _ptr(DEPTH_Q), C.c_uint64(seed),
                _ptr(p), _ptr(vi), _ptr(a), _ptr(d), _ptr(labels))
    return p, vi, a, d, labels
