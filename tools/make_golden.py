"""Writes tests/golden/oracle_example_donor.npz: golden vectors of the hot path on the reference's own fixture
(data/example_donor.RData, decoded into tests/golden/example_donor.npz by tools/rdata_to_npz.py).

The reference is R and cannot run here (parity unpinned), so the vectors come from the NumPy
restatement in oracle/: a recorded Gibbs chain (every uniform of sample() / rbinom(), every Beta draw, and the
resulting assignments, Config flips, probabilities and logLik trace), the EM fixed point from a stored start, and
get_logLik at a stored state.  tests/test_oracle.py replays the recording through the oracle (pins the oracle
against drift); tests/test_gpu_parity.py replays it through the CUDA library (cdl_replay) and through cdl_em_run /
cdl_get_loglik.  A genuine R recording made with tools/record_reference.R has the same fields and can replace it.

    python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cardelino_oracle as O  # noqa: E402

T, SEED = 30, 20261019


def main():
    d = np.load(os.path.join(ROOT, "tests", "golden", "example_donor.npz"))
    A, D, Z = d["A"], d["D"], d["Z"]
    N, M = A.shape
    K = Z.shape[1]
    rec = O.RecordingDraws(SEED, T, M, N * K)
    o = O.clone_id_Gibbs(A, D, Z, min_iter=T, max_iter=T, draws=rec)
    theta_init = np.array([0.02, 0.6])
    em = O.clone_id_EM(A, D, Z, theta_init=theta_init)
    A1, B1, _ = O.preprocess(A, D)
    it = 17                                                    # get_logLik at the state after sweep 17
    cfg_it = o["Config_all"][it - 1].reshape(K, N).T
    ll_state = O.get_logLik_tables(A1, B1, cfg_it, o["assign_all"][it - 1], o["theta0_all"].ravel()[it - 1],
                                   o["theta1_all"][it - 1])
    out = os.path.join(ROOT, "tests", "golden", "oracle_example_donor.npz")
    np.savez_compressed(
        out, T=T, u_assign=rec.ua, u_config=rec.uc, theta0_all=o["theta0_all"].ravel(), theta1_all=o["theta1_all"],
        relax_rate_all=o["relax_rate_all"], assign_all=o["assign_all"].astype(np.int16),
        Config_all=o["Config_all"].astype(np.uint8), prob=o["prob"], Config_prob=o["Config_prob"],
        logLik=np.asarray(o["logLik"], dtype=np.float64), prob_all_last=o["prob_all"][T - 1],
        em_theta_init=theta_init, em_theta=np.asarray(em["theta"], dtype=np.float64), em_logLik=float(em["logLik"]),
        em_prob=em["prob"], em_n_iter=int(em.get("n_iter", -1)),
        ll_it=it, ll_value=ll_state)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
