"""Where a clone_id(n_chain = 64) call on example_donor (bench.py --workload C5) spends its host time: cProfile over the
Python side (ctypes calls show up as the _lib.Handle methods) and, with CDL_PHASES=1, the library's own phase timers."""
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CDL_PHASES", "1")
import cardelino_b200 as cb  # noqa: E402

d = np.load(os.path.join(ROOT, "tests", "golden", "example_donor.npz"))
A, D, Z = d["A"], d["D"], d["Z"]
h = cb.Handle(0)
kw = dict(min_iter=2000, max_iter=5000, seed=20261020, handle=h, verbose=False, n_chain=64, light=True)
import warnings
warnings.simplefilter("ignore")
cb.clone_id(A, D, Config=Z, **kw)
t0 = time.perf_counter()
cb.clone_id(A, D, Config=Z, **kw)
print("one call: %.3f s" % (time.perf_counter() - t0))
pr = cProfile.Profile()
pr.enable()
res = cb.clone_id(A, D, Config=Z, **kw)
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18)
print(s.getvalue())
