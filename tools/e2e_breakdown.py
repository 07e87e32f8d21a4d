import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import cardelino_b200 as cb, bench
N, M, K, cov = bench.WORKLOADS["C4"]
cfg = bench.make_config(N, K)
h = cb.Handle(0)
h.synth_generate(N, M, K, cfg, cov, bench.SEED_DATA)
pin = lambda x: torch.from_numpy(x.view(np.uint8)).pin_memory().numpy().view(x.dtype)
p, i, a, d = [pin(x) for x in h.export_csc_u16()]
def t(label, f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); print('%-28s %.3f s' % (label, time.perf_counter() - t0)); return r
for rep in range(2):
    t('load_csc_u16 (H2D + build)', lambda: h.load_csc_u16(N, M, p, i, a, d))
    t('gibbs_run 1000 sweeps', lambda: h.gibbs_run(cfg, None, min_iter=1000, max_iter=1000, seed=1, keep_prob_all=0))
    t('fetch prob', lambda: h.fetch("prob"))
    t('fetch Config_prob', lambda: h.fetch("Config_prob"))
    t('fetch theta_traces', lambda: h.fetch("theta_traces"))
    t('fetch theta', lambda: h.fetch("theta"))
    t('fetch logLik/relax', lambda: (h.fetch("logLik"), h.fetch("relax_rate")))
