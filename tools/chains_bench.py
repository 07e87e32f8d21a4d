"""Throughput of many independent chains on a SMALL donor (the reference's own use case: example_donor,
34 variants x 428 cells x 4 clones; BASELINE config 5 runs 64 chains).  Prints chain-iterations/s for n_chain in
{1, 4, 16, 64}: chains are interleaved on up to 4 streams inside the library."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import cardelino_b200 as cb
    d = np.load(os.path.join(ROOT, "tests", "golden", "example_donor.npz"))
    A, D, Z = d["A"], d["D"], d["Z"]
    h = cb.Handle(0)
    h.load_dense(A, D)
    T = 600
    out = {}
    for n in (1, 4, 16, 64):
        h.gibbs_run(Z, None, min_iter=T, max_iter=T, seed=3, n_chain=n, keep_prob_all=1)      # warm
        t0 = time.perf_counter()
        h.gibbs_run(Z, None, min_iter=T, max_iter=T, seed=3, n_chain=n, keep_prob_all=1)
        dt = time.perf_counter() - t0
        out[n] = {"seconds": dt, "chain_iterations_per_s": n * (T - 1) / dt}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
