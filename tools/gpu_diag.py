"""First-contact GPU diagnostic: runs each entry point on small inputs, compares with the oracle and
prints everything (used under gpurun while bringing the kernels up)."""
import sys, time, traceback
import numpy as np
sys.path.insert(0, ".")
import cardelino_b200 as cb
from oracle import cardelino_oracle as O

d = np.load("tests/golden/example_donor.npz")
A, D, Z = d["A"], d["D"], d["Z"]
h = cb.Handle(0)

def step(name, f):
    t = time.time()
    try:
        f()
        print("[ok] %s (%.2fs)" % (name, time.time() - t), flush=True)
    except Exception:
        print("[FAIL] %s" % name, flush=True)
        traceback.print_exc()

def t_load():
    h.load_dense(A, D)
    print(" info", h.data_info())
    A1, B1, W = O.preprocess(A, D)
    print(" W_log oracle", W)
    p, i, a, dd = h.export_csc_u16()
    assert p[-1] == 1874, p[-1]
    dense_d = np.zeros_like(A1); dense_a = np.zeros_like(A1)
    for j in range(A.shape[1]):
        dense_d[i[p[j]:p[j+1]], j] = dd[p[j]:p[j+1]]; dense_a[i[p[j]:p[j+1]], j] = a[p[j]:p[j+1]]
    assert np.array_equal(dense_a, A1) and np.array_equal(dense_d, A1 + B1)

def t_em():
    r = h.em_run(Z, theta_init=[0.1, 0.5])
    o = O.clone_id_EM(A, D, Z, theta_init=[0.1, 0.5])
    print(" gpu", r["theta"], r["logLik"], r["n_iter"], " oracle", o["theta"], o["logLik"], o["n_iter"])
    print(" max|dprob|", np.abs(r["prob"] - o["prob"]).max())

def t_ll():
    A1, B1, W = O.preprocess(A, D)
    rng = np.random.default_rng(0)
    lab = rng.integers(1, 5, A.shape[1])
    th1 = rng.uniform(0.2, 0.8, A.shape[0])
    g = h.get_loglik(Z, lab, 0.03, th1)
    o = O.get_logLik(A1, B1, Z, lab, 0.03, np.repeat(th1.reshape(-1, 1), A.shape[1], 1))
    print(" labels gpu", g, "oracle", o)
    P = rng.dirichlet(np.ones(4), A.shape[1]); C = rng.uniform(0, 1, Z.shape)
    g = h.get_loglik(C, P, 0.03, th1)
    o = O.get_logLik(A1, B1, C, P, 0.03, np.repeat(th1.reshape(-1, 1), A.shape[1], 1))
    print(" fractional gpu", g, "oracle", o)

def t_gibbs(**kw):
    T = kw.pop("T", 60)
    cfg = kw.pop("Config", Z)
    res = cb.clone_id_Gibbs(A, D, cfg, min_iter=T, max_iter=T, seed=11, keep_assign_all=True, handle=h,
                            verbose=False, **kw)
    dr = O.ReplayDraws(seed=11, chain=0, theta0_all=res["theta0_all"], theta1_all=res["theta1_all"],
                       relax_rate_all=res["relax_rate_all"])
    okw = dict(kw)
    o = O.clone_id_Gibbs(A, D, cfg, min_iter=T, max_iter=T, draws=dr, **okw)
    print("  n_iter", res["n_iter"], o["n_iter"], "theta0", res["theta0"], o["theta0"])
    print("  max|dprob_all|", np.abs(res["prob_all"] - o["prob_all"]).max())
    mism = (res["assign_all"] != o["assign_all"]).sum()
    print("  assign mismatches", mism, "of", res["assign_all"].size)
    print("  Config_all mismatches", (res["Config_all"] != o["Config_all"]).sum())
    print("  max|dlogLik|", np.abs(res["logLik"] - o["logLik"]).max(), "logLik[-1]", res["logLik"][-1], o["logLik"][-1])
    print("  max|dprob|", np.abs(res["prob"] - o["prob"]).max(), "max|dConfig_prob|", np.abs(res["Config_prob"] - o["Config_prob"]).max())
    print("  relax_rate", res["relax_rate"], o["relax_rate"], "DIC", res["DIC"]["DIC"], o["DIC"]["DIC"])
    print("  max|dprob_variant|", np.nanmax(np.abs(res["prob_variant"] - o["prob_variant"])))
    print("  pc", res["percent_converged"], o["percent_converged"], "logLik_post", res["logLik_post"], o["logLik_post"])
    v = O.verify_gibbs_trace(A, D, cfg, res, min_iter=T, seed=11, **{k_: v_ for k_, v_ in kw.items() if k_ in ("relax_Config", "relax_rate_fixed", "wise")})
    print("  VERIFY", v)

step("load", t_load)
step("em", t_em)
step("get_loglik", t_ll)
step("gibbs variant/learned", lambda: t_gibbs())
step("gibbs global/fixed", lambda: t_gibbs(wise="global", relax_rate_fixed=0.3))
step("gibbs relax off", lambda: t_gibbs(relax_Config=False))
step("gibbs de-novo", lambda: t_gibbs(Config=np.zeros((34, 3)), relax_rate_fixed=0.5, T=40))
step("gibbs de-novo K=6", lambda: t_gibbs(Config=np.zeros((34, 6)), relax_rate_fixed=0.5, T=40))
def t_long():
    t = time.time()
    res = cb.clone_id_Gibbs(A, D, Z, min_iter=800, max_iter=1200, seed=3, handle=h, verbose=True)
    print("  long run", time.time() - t, "s; n_iter", res["n_iter"], res["percent_converged"], res["theta0"], res["relax_rate"])
step("gibbs long", t_long)
def t_synth():
    rng = np.random.default_rng(1)
    N, M, K = 2000, 100000, 8
    cfg = (rng.random((N, K)) < 0.3).astype(float); cfg[:, 0] = 0
    t = time.time(); h.synth_generate(N, M, K, cfg, 0.19, 5); print("  synth", time.time() - t, h.data_info())
    t = time.time()
    h.gibbs_run(cfg, None, min_iter=20, max_iter=20, seed=1, keep_prob_all=1)
    print("  run 20 its", time.time() - t)
    print("  bench", h.bench_sweeps(3, 20))
    lab = h.synth_labels(); pr = h.fetch("prob"); print("  acc", (pr.argmax(1) + 1 == lab).mean())
step("synth C3", t_synth)

def t_synth4():
    rng = np.random.default_rng(2)
    N, M, K = 10000, 1000000, 16
    cfg = (rng.random((N, K)) < 0.3).astype(float); cfg[:, 0] = 0
    t = time.time(); h.synth_generate(N, M, K, cfg, 0.05, 7); print("  synth", time.time() - t, h.data_info())
    t = time.time()
    h.gibbs_run(cfg, None, min_iter=12, max_iter=12, seed=1, keep_prob_all=1)
    print("  run 12 its", time.time() - t)
    print("  bench", h.bench_sweeps(3, 10))
    lab = h.synth_labels(); pr = h.fetch("prob"); print("  acc", (pr.argmax(1) + 1 == lab).mean())
step("synth C4", t_synth4)
