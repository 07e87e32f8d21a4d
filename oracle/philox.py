"""TEST INFRASTRUCTURE ONLY (checker) -- never imported by the product path.

Counter-based Philox4x32-10 (Salmon et al., SC'11, Random123) restated in NumPy so the oracle can
reproduce, bit for bit, the uniforms the CUDA sampler consumes.  The reference (pure R, Mersenne
Twister through `sample`/`rbinom`, R/clone_id.R:493,534) has no counter-based generator; this is the
B200 build's own keying so that draws are identical at any GPU count:

    key     = (seed_lo, seed_hi)
    counter = (idx, sub, iteration, (chain << 8) | stream)

`idx` is the global cell id (stream ASSIGN) or the column-major Config index i + k*N (stream CONFIG);
`sub` is 0 for those two streams (rejection samplers for Beta draws bump it per attempt).
The same constants live in cardelino_b200/csrc/philox.cuh.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)

STREAM_ASSIGN = 0
STREAM_CONFIG = 1
STREAM_THETA0 = 2
STREAM_THETA1 = 3
STREAM_RELAX = 4
STREAM_EM_INIT = 5


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=10):
    """Vectorised Philox4x32; all inputs broadcastable to a common shape, dtype uint32."""
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint32) for x in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    c0, c1, c2, c3 = (x.astype(np.uint64) for x in (c0, c1, c2, c3))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(rounds):
            p0 = M0 * c0
            p1 = M1 * c2
            hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
            hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
            n0 = (hi1 ^ c1 ^ np.uint64(k0)) & MASK
            n1 = lo1
            n2 = (hi0 ^ c3 ^ np.uint64(k1)) & MASK
            n3 = lo0
            c0, c1, c2, c3 = n0, n1, n2, n3
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return tuple(x.astype(np.uint32) for x in (c0, c1, c2, c3))


def u53(x0, x1):
    """Two 32-bit words -> double strictly inside (0, 1) with 53-bit resolution."""
    hi = (np.asarray(x0, dtype=np.uint64) >> np.uint64(5)).astype(np.float64)   # 27 bits
    lo = (np.asarray(x1, dtype=np.uint64) >> np.uint64(6)).astype(np.float64)   # 26 bits
    return (hi * 67108864.0 + lo + 0.5) * (1.0 / 9007199254740992.0)


def uniforms(seed, chain, iteration, stream, idx, sub=0):
    """Uniform(0,1) doubles for element indices `idx` (array) of one (chain, iteration, stream)."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    k0, k1 = seed & 0xFFFFFFFF, seed >> 32
    idx = np.asarray(idx, dtype=np.uint32)
    c3 = np.uint32(((int(chain) << 8) | int(stream)) & 0xFFFFFFFF)
    x0, x1, _, _ = philox4x32(idx, np.uint32(sub), np.uint32(iteration), c3, k0, k1)
    return u53(x0, x1)


def uniform_pairs(seed, chain, iteration, stream, idx, sub):
    """The two 53-bit uniforms of one Philox block (PhiloxSeq::next2 in csrc/common.cuh); idx / sub broadcast."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    k0, k1 = seed & 0xFFFFFFFF, seed >> 32
    c3 = np.uint32(((int(chain) << 8) | int(stream)) & 0xFFFFFFFF)
    x0, x1, x2, x3 = philox4x32(np.asarray(idx, np.uint32), np.asarray(sub, np.uint32), np.uint32(iteration), c3, k0, k1)
    return u53(x0, x1), u53(x2, x3)


def gamma_mt(shape, seed, chain, iteration, stream, idx, sub0=0):
    """Marsaglia-Tsang Gamma(shape, 1) exactly as csrc/common.cuh::gamma_mt walks its Philox sub-stream (vectorised
    over `shape` / `idx`): used to check the Beta draws of the CUDA table kernel against the shape parameters
    prior + S (R/clone_id.R:563-571).  Device libm differs from NumPy's in the last bits, so draws agree to
    rounding, not bitwise (and a rejection test can, very rarely, fall the other way)."""
    a = np.array(shape, dtype=np.float64, ndmin=1).copy()
    idx = np.broadcast_to(np.asarray(idx, np.uint32), a.shape).copy()
    sub = np.full(a.shape, int(sub0), dtype=np.uint64)
    boost = np.ones_like(a)
    small = a < 1.0
    if small.any():
        u, _ = uniform_pairs(seed, chain, iteration, stream, idx[small], sub[small].astype(np.uint32))
        boost[small] = u ** (1.0 / a[small])
        a[small] += 1.0
        sub[small] += 1
    d = a - 1.0 / 3.0
    c = 1.0 / np.sqrt(9.0 * d)
    out = np.full(a.shape, np.nan)
    todo = np.ones(a.shape, dtype=bool)
    for _ in range(1000):
        if not todo.any():
            break
        t = np.flatnonzero(todo)
        u1, u2 = uniform_pairs(seed, chain, iteration, stream, idx[t], sub[t].astype(np.uint32))
        u3, _ = uniform_pairs(seed, chain, iteration, stream, idx[t], (sub[t] + 1).astype(np.uint32))
        sub[t] += 2
        x = np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
        v = 1.0 + c[t] * x
        ok = v > 0.0
        v3 = np.where(ok, v, 1.0) ** 3
        x2 = x * x
        with np.errstate(invalid="ignore", divide="ignore"):
            acc = ok & ((u3 < 1.0 - 0.0331 * x2 * x2) | (np.log(u3) < 0.5 * x2 + d[t] * (1.0 - v3 + np.log(v3))))
        out[t[acc]] = boost[t[acc]] * (d[t[acc]] * v3[acc])     # gamma_mt: boost * gamma_mt_attempts(...) = boost * (d * v)
        todo[t[acc]] = False
    return out


def _clamp_theta(t):
    return np.clip(t, 1e-300, 1.0 - 1.1102230246251565e-16)


def beta_parallel(a, b, seed, chain, iteration, stream, idx):
    """Beta(a, b) as the sweep's table kernel draws it: X ~ Gamma(a) on sub-stream 0, Y ~ Gamma(b) on sub-stream
    2^31 of the same (chain, iteration, stream, idx) -- two jobs on two threads (csrc/gibbs.cu::k_table)."""
    x = gamma_mt(a, seed, chain, iteration, stream, idx, 0)
    y = gamma_mt(b, seed, chain, iteration, stream, idx, 1 << 31)
    s = x + y
    return _clamp_theta(np.where(s > 0.0, x / np.where(s > 0.0, s, 1.0), 0.5))
