"""cProfile of the end-to-end clone_id_Gibbs call of bench.py (host buffers, C4), to see where the time outside
the sweeps goes.  Usage: python tools/e2e_profile.py [T]"""
import sys, time, cProfile, pstats, numpy as np, torch
sys.path.insert(0, '/root/repo')
import cardelino_b200 as cb, bench
T = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
N, M, K, cov = bench.WORKLOADS["C4"]
cfg = bench.make_config(N, K)
h = cb.Handle(0)
h.synth_generate(N, M, K, cfg, cov, bench.SEED_DATA)
pin = lambda x: torch.from_numpy(x.view(np.uint8)).pin_memory().numpy().view(x.dtype)
p, i, a, d = [pin(x) for x in h.export_csc_u16()]
counts = cb.SparseCounts(N, M, p, i, a, d)
def run():
    r = cb.clone_id_Gibbs(counts, None, cfg, min_iter=T, max_iter=T, seed=1, verbose=False, handle=h,
                          keep_prob_all=0, light=True)
    torch.cuda.synchronize(); return r
run()
t0 = time.perf_counter(); run(); print("second call: %.3f s" % (time.perf_counter() - t0))
pr = cProfile.Profile(); pr.enable(); run(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
