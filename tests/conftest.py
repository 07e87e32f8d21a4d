import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def example_donor():
    d = np.load(os.path.join(GOLDEN, "example_donor.npz"), allow_pickle=False)
    return d["A"], d["D"], d["Z"]


@pytest.fixture(scope="session")
def simulation_input():
    d = np.load(os.path.join(GOLDEN, "simulation_input.npz"))
    return d["D_input"], d["tree_4clone_Z"]


def sim_read_count(Config, D, rng, cell_num, means=(0.002, 0.45), vars_=(100, 1)):
    """NumPy restatement of R/reads_simulator.R:53-152 (wise0 = element, wise1 = variant, no doublets);
    test-input generator only."""
    N, K = Config.shape
    D = np.array(D, float)
    D[D == 0] = np.nan
    D_sim = D[:, rng.integers(0, D.shape[1], cell_num)]
    lab = rng.integers(0, K, cell_num)
    H = Config[:, lab]
    th1 = rng.beta(means[1] * vars_[1], (1 - means[1]) * vars_[1], N)
    th0 = rng.beta(means[0] * vars_[0], (1 - means[0]) * vars_[0], (N, cell_num))
    p = np.where(H > 0, th1[:, None], th0)
    cov = ~np.isnan(D_sim)
    A = np.full(D_sim.shape, np.nan)
    A[cov] = rng.binomial(D_sim[cov].astype(int), p[cov])
    return A, D_sim, lab + 1


def sample_seq_depth(D, rng, n_cells, n_sites):
    """R/reads_simulator.R:174-212 with missing_rate = NULL."""
    D = np.array(D, float)
    D[D == 0] = np.nan
    miss = np.isnan(D).mean(axis=1)
    idx = rng.choice(D.shape[0], n_sites, replace=n_sites > D.shape[0])
    out = np.full((n_sites, n_cells), np.nan)
    for r, i in enumerate(idx):
        vals = D[i][~np.isnan(D[i])]
        if vals.size == 0:
            continue
        out[r] = rng.choice(vals, n_cells, replace=True)
        out[r, rng.choice(n_cells, int(round(miss[i] * n_cells)), replace=False)] = np.nan
    return out


@pytest.fixture(scope="session")
def c2_donor(simulation_input):
    """BASELINE config 2: sim_read_count(tree_4clone$Z, sample_seq_depth(D_input, 500, 403), cell_num=500)."""
    D_input, Z = simulation_input
    rng = np.random.default_rng(20261019 + 2)
    D2 = sample_seq_depth(D_input, rng, 500, Z.shape[0])
    A, D, lab = sim_read_count(Z, D2, rng, 500)
    return A, D, Z, lab


@pytest.fixture(scope="session")
def gpu_handle():
    import cardelino_b200 as cb
    h = cb.Handle(0)
    yield h
    h.close()
