"""The R side of the boundary, executed without R: r/cardelino_b200_shim.c is compiled against a small stand-in for R's C
API (tests/r_stub/fake_r.c: tagged vectors with names / dim, PROTECT counting, Rf_error as a longjmp to the caller's
context, R_ToplevelExec, an interrupt that can be armed) and its .Call entry points are driven by
tests/r_stub/shim_driver.c exactly as R would call them.  The CPU test covers registration and the error path of
cdl_create; the GPU test runs every entry point and compares what the shim hands to R with the ctypes path."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_driver(tmp_path):
    exe = str(tmp_path / "shim_driver")
    libdir = os.path.join(ROOT, "cardelino_b200")
    cmd = ["gcc", "-Wall", "-Werror", "-O1", "-I", os.path.join(ROOT, "tests", "r_stub"), "-I", os.path.join(ROOT, "include"),
           "-o", exe, os.path.join(ROOT, "r", "cardelino_b200_shim.c"), os.path.join(ROOT, "tests", "r_stub", "fake_r.c"),
           os.path.join(ROOT, "tests", "r_stub", "shim_driver.c"), "-L", libdir, "-lcardelino_b200",
           "-Wl,-rpath," + libdir, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_shim_registration_and_error_paths(tmp_path):
    """No GPU needed: the registration table answers .Call lookups with R's argument-count check, a released handle
    raises an R error, and cdl_create's failure on a box without a B200 arrives as an R error -- no crash, no CPU
    fallback (on a GPU box the handle is created and finalized instead)."""
    import cardelino_b200 as cb
    if not os.path.exists(cb.LIB_PATH):
        from cardelino_b200 import build
        build.build()
    exe = _build_driver(tmp_path)
    r = subprocess.run([exe, "errors"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ERRORS MODE OK" in r.stdout and "lookup checks: ok" in r.stdout and "released-handle check: ok" in r.stdout


def _parse(path):
    out = {}
    with open(path) as f:
        for line in f:
            name, n, rest = line.rstrip("\n").split(" ", 2) if line.count(" ") >= 2 else (line.split()[0], line.split()[1], "")
            n = int(n)
            out[name] = rest if n < 0 else np.array([float(x) for x in rest.split()], dtype=np.float64)
    return out


@pytest.mark.gpu
def test_shim_entry_points_match_ctypes_path(tmp_path, example_donor, gpu_handle):
    """Every .Call entry point on a B200 (load dense / dgCMatrix, Gibbs with the full list, the light list, relabel,
    two chains, verbose printing through Rprintf, EM, get_logLik, the finalizer), its error paths (each an R error
    with the reference's wording, handle still usable afterwards), and a user interrupt in mid-run.  The numbers R
    would receive equal the ctypes path's on the same seed."""
    import cardelino_b200 as cb
    A, D, Z = example_donor
    N, M = A.shape
    K = Z.shape[1]
    T0, T1, seed = 200, 500, 77                    # the reference's own test call (tests/testthat/test-clone_id.R:16-20)
    inp = tmp_path / "in.bin"
    with open(inp, "wb") as f:
        np.array([N, M, K, T0, T1, seed], np.int64).tofile(f)
        for X in (A, D, Z):
            np.asfortranarray(X, dtype=np.float64).ravel(order="F").tofile(f)
    exe = _build_driver(tmp_path)
    outp = tmp_path / "out.txt"
    r = subprocess.run([exe, "run", str(inp), str(outp)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "RUN MODE OK" in r.stdout, r.stdout + r.stderr
    o = _parse(outp)
    kw = dict(min_iter=T0, max_iter=T1, seed=seed, handle=gpu_handle, verbose=False)
    full = cb.clone_id_Gibbs(A, D, Z, n_chain=2, keep_prob_all=1, _all_chains=True, **kw)
    for c in range(2):
        g, pre = full[c], "full%d." % c
        it = int(o[pre + "n_iter"][0])
        assert it == g["n_iter"] and o[pre + "percent_converged"][0] == g["percent_converged"]
        assert np.array_equal(o[pre + "prob"].reshape(K, M).T, g["prob"])
        assert np.array_equal(o[pre + "Config_prob"].reshape(K, N).T, g["Config_prob"])
        assert np.array_equal(o[pre + "prob_all_t"].reshape(it, M * K), g["prob_all"])          # (M*K) x it column-major
        assert np.array_equal(o[pre + "Config_all_t"].reshape(it, N * K), g["Config_all"])
        assert np.array_equal(o[pre + "prob_variant"].reshape(M, N).T, g["prob_variant"])
        assert np.array_equal(o[pre + "theta1_all"].reshape(N, it).T, g["theta1_all"])
        assert np.array_equal(o[pre + "logLik"], g["logLik"]) and o[pre + "theta0"][0] == g["theta0"]
        assert o[pre + "relax_rate"][0] == g["relax_rate"] and np.array_equal(o[pre + "relax_rate_all"], g["relax_rate_all"])
        assert np.isclose(o[pre + "DIC"][0], g["DIC"]["DIC"], rtol=1e-12) and o[pre + "stop_rule"][0] == 0
        assert o[pre + "logLik_post"][0] == g["logLik_post"]
    # verbose: one "x% converged." line per convergence check of each chain (R/clone_id.R:583-588)
    lines = [x for x in o["full.stdout"].split("|") if x]
    checks = sum(len(range(T0, int(o["full%d.n_iter" % c][0]) + 1, 100)) for c in range(2))
    assert len(lines) == checks and all(x.endswith("% converged.") for x in lines), lines
    # light list: no big arrays, the stop rule from running sums, same chain
    lg = cb.clone_id_Gibbs(A, D, Z, keep_prob_all=0, light=True, **kw)
    assert o["light.prob_all_t"].size == 0 and o["light.Config_all_t"].size == 0 and o["light.prob_variant"].size == 0
    assert int(o["light.n_iter"][0]) == lg["n_iter"] and o["light.stop_rule"][0] == lg["stop_rule"] and o["light.stdout"] == ""
    assert np.array_equal(o["light.prob"].reshape(K, M).T, lg["prob"])
    assert np.array_equal(o["light.theta1_elements"], lg["theta1_elements"])
    # relabel (R/clone_id.R:624-644) inside the library
    rl = cb.clone_id_Gibbs(A, D, Z, relabel=True, **kw)
    assert np.array_equal(o["relabel.prob"].reshape(K, M).T, rl["prob"])
    assert np.array_equal(o["relabel.prob_all_t"].reshape(rl["n_iter"], M * K), rl["prob_all"])
    # error paths, with the reference's wording where it has one (R/clone_id.R:368-373,437-439,422-424)
    assert "relabel needs the prob_all trace" in o["err.relabel_light"]
    assert o["err.wise"] == "Input wise mode: foo, while only supporting: element, variant, global"
    assert o["err.relax_rate_fixed"] == "relax_rate_fixed needs to be NULL or in [0, 1]."
    assert "prior1" in o["err.prior1_rows"] and "unknown parameter 'bogus'" in o["err.unknown_param"]
    assert o["err.interrupt"] == "interrupted by the user" and o["interrupt.polls"][0] == 3
    assert o["err.released"] == "cardelino_b200: handle already released"
    assert o["protect"][0] == 0 and o["protect"][1] >= 10          # balanced PROTECT / UNPROTECT
    # EM and get_logLik
    em = cb.clone_id_EM(A, D, Z, theta_init=[0.1, 0.5], handle=gpu_handle, verbose=False)
    assert np.array_equal(o["em.theta"], em["theta"]) and o["em.logLik"][0] == em["logLik"] and o["em.n_iter"][0] == em["n_iter"]
    assert np.array_equal(o["em.prob"].reshape(K, M).T, em["prob"])
    th1 = 0.3 + 0.4 * np.arange(N) / N
    assert o["loglik.labels"][0] == gpu_handle.get_loglik(Z, 1 + np.arange(M) % K, 0.02, th1)
    assert o["loglik.prob"][0] == gpu_handle.get_loglik(Z, em["prob"], 0.02, th1)
    # dgCMatrix loader gives the dense loader's donor
    assert np.array_equal(o["dgc.prob"].reshape(K, M).T, lg["prob"]) and np.array_equal(o["dgc.logLik"], lg["logLik"])
