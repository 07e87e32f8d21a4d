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
