"""CPU tests: the C-ABI library loads and exports every symbol the header declares, fails loudly without
a GPU (no CPU fallback), and the host mirror reproduces the reference's argument checks."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    import cardelino_b200 as cb
    if not os.path.exists(cb.LIB_PATH):
        g.build()
    return ctypes.CDLL(cb.LIB_PATH)


def test_header_symbols_exported(lib):
    import cardelino_b200 as cb
    hdr = open(os.path.join(ROOT, "include", "cardelino_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(cdl_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), "missing export " + name
    assert sorted(cb.EXPORTS) == declared


def test_struct_layout_matches_defaults(lib):
    from cardelino_b200._lib import GibbsParams
    p = GibbsParams()
    lib.cdl_gibbs_defaults.restype = None
    lib.cdl_gibbs_defaults(ctypes.byref(p))
    assert (p.relax_Config, p.keep_base_clone, p.min_iter, p.max_iter, p.wise, p.n_chain) == (1, 1, 5000, 20000, 1, 1)
    assert list(p.relax_rate_prior) == [1, 9] and list(p.prior0) == [0.2, 99.8] and list(p.prior1) == [0.45, 0.55]
    assert p.buin_frac == 0.5 and p.check_every == 100 and p.keep_prob_all == -1


def test_no_cpu_fallback():
    import torch
    import cardelino_b200 as cb
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(cb.CdlError) as e:
        cb.Handle(0)
    assert e.value.code == 2 and "no CPU path" in str(e.value)


def test_product_never_imports_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "cardelino_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_argument_checks_match_reference_messages(example_donor):
    import cardelino_b200 as cb
    A, D, Z = example_donor
    with pytest.raises(ValueError, match="Config and n_clone can't be NULL together"):
        cb.clone_id(A, D)
    with pytest.raises(ValueError, match="non-integer"):
        cb.clone_id(np.where(np.isnan(A), np.nan, A + 0.25), D, Z)
    with pytest.raises(ValueError, match="should be one of"):
        cb.clone_id(A, D, Z, inference="vb")
    with pytest.raises(ValueError, match="only supporting: element, variant, global"):
        cb.clone_id_Gibbs(A, D, Z, wise="cell")
    with pytest.raises(ValueError, match=r"relax_rate_fixed needs to be NULL or in \[0, 1\]"):
        cb.clone_id_Gibbs(A, D, Z, relax_rate_fixed=-0.1)
    with pytest.raises(ValueError, match="must have the same size"):
        cb.clone_id_Gibbs(A, D[:, :-1], Z)
    with pytest.raises(ValueError, match="must have the same size"):
        cb.clone_id_EM(A, D, Z[:-1])
    with pytest.raises(ValueError, match="prior1 need to be a matrix"):
        cb.clone_id_Gibbs(A, D, Z, prior1=(1, 2, 3))


def test_dimnames_matching(example_donor):
    import pandas as pd
    import cardelino_b200 as cb
    A, D, Z = example_donor
    var = ["v%d" % i for i in range(34)]
    cells = ["c%d" % j for j in range(428)]
    Adf, Ddf = pd.DataFrame(A, index=var, columns=cells), pd.DataFrame(D, index=var, columns=cells)
    with pytest.raises(ValueError, match="Rownames for A and D are not identical"):
        cb.clone_id(Adf, pd.DataFrame(D, index=var[::-1], columns=cells), Z)
    with pytest.raises(ValueError, match="No matches in variant names"):
        cb.clone_id(Adf, Ddf, pd.DataFrame(Z, index=["w%d" % i for i in range(34)]))


def test_colmatch_and_helpers():
    import cardelino_b200 as cb
    rng = np.random.default_rng(0)
    for K in (3, 5, 10):
        A = rng.random((40, K))
        perm = rng.permutation(K)
        B = A[:, perm] + 0.005 * rng.random((40, K))
        for force in (True, False):
            assert np.array_equal(perm[cb.colMatch(A, B, force=force)], np.arange(K))
    X = rng.random((300, 5))
    from oracle import cardelino_oracle as O
    assert np.allclose(cb.Geweke_Z(X), O.geweke_z(X))
    d = cb.devianceIC([-10.0, -12.0, -11.0], -9.0, verbose=False)
    assert d["DIC"] == pytest.approx(22.0)
    out = cb.assign_cells_to_clones(np.array([[0.7, 0.3], [0.5, 0.5]]))
    assert out["clone"] == ["clone1", "unassigned"]


def test_bench_reference_arm_line():
    """bench.py --impl reference prints the contract's JSON line (tiny sample so the CPU suite stays fast)."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny",
                        "--steps", "1", "--warmup", "0", "--cpu-cells", "16", "--cpu-iters", "1"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0


def test_bench_reference_arm_under_the_launcher():
    """The driver starts the reference arm like ours: under torch.distributed.run for N > 1, which exports
    OMP_NUM_THREADS=1.  Rank 0 alone works and prints ONE line, on all host threads (round 1 inherited the launcher's
    single thread); the other ranks exit 0 without work."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29561", os.path.join(ROOT, "bench.py"),
                        "--impl", "reference", "--gpus", "2", "--workload", "tiny", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["steps"] == 2 and line["warmup"] == 1
    assert line["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["d2h_bytes_per_step"] == 0


def test_r_shim_compiles_against_stub_headers():
    """The .Call shim cannot be built here (no R), but it must at least be valid C against the C ABI header: compile
    it with -fsyntax-only -Wall -Werror against stand-in declarations of the R API entry points it uses."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(root, "tests", "r_stub"),
                        "-I", os.path.join(root, "include"), os.path.join(root, "r", "cardelino_b200_shim.c")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_sparse_counts_from_sparse_and_mtx(tmp_path):
    """Sparse ingestion without densifying: two sparse matrices / MatrixMarket files
    -> the one-pattern uint16 form, with the reference's missing-data rules (R/clone_id.R:380-386) and the
    library's limits."""
    import scipy.sparse as sp
    from scipy.io import mmwrite
    import cardelino_b200 as cb
    d = np.load(os.path.join(ROOT, "tests", "golden", "mito_passed.npz"))
    AD, DP = d["AD"], d["DP"]                       # the reference's inst/extdata/passed_{ad,dp}.mtx, decoded
    rng = np.random.default_rng(0)
    DP = DP * (rng.random(DP.shape) < 0.7)          # knock out a third of the entries: D == 0 is missing
    AD = np.minimum(AD, 60000)
    DP = np.maximum(np.minimum(DP, 65535), AD) * (DP > 0)
    stray = AD * (DP == 0)                          # A entries where D is missing must be dropped
    assert stray.sum() > 0
    sc = cb.SparseCounts.from_sparse(sp.coo_matrix(AD), sp.csr_matrix(DP))
    A2, D2 = sc.to_dense()
    assert np.array_equal(np.isnan(D2), DP == 0) and np.array_equal(np.isnan(A2), DP == 0)
    assert np.array_equal(np.nan_to_num(D2), DP) and np.array_equal(np.nan_to_num(A2), AD * (DP > 0))
    assert sc.a.dtype == np.uint16 and sc.d.dtype == np.uint16 and sc.indptr[-1] == (DP > 0).sum()
    mmwrite(str(tmp_path / "ad.mtx"), sp.coo_matrix(AD))
    mmwrite(str(tmp_path / "dp.mtx"), sp.coo_matrix(DP))
    sm = cb.SparseCounts.from_mtx(str(tmp_path / "ad.mtx"), str(tmp_path / "dp.mtx"))
    assert np.array_equal(sm.indptr, sc.indptr) and np.array_equal(sm.indices, sc.indices)
    assert np.array_equal(sm.a, sc.a) and np.array_equal(sm.d, sc.d)
    with pytest.raises(ValueError, match="non-integer"):
        cb.SparseCounts.from_sparse(sp.csc_matrix(AD * 0.5), sp.csc_matrix(DP))
    with pytest.raises(ValueError, match="65535"):
        cb.SparseCounts.from_sparse(sp.csc_matrix(AD), sp.csc_matrix(DP * 100))
    with pytest.raises(ValueError, match="exceeds"):
        cb.SparseCounts.from_sparse(sp.csc_matrix(DP + (DP > 0)), sp.csc_matrix(DP))
    with pytest.raises(ValueError, match="same size"):
        cb.SparseCounts.from_sparse(sp.csc_matrix(AD[:-1]), sp.csc_matrix(DP))


def test_clone_id_accepts_compact_counts_up_to_the_device():
    """clone_id(SparseCounts, None, ...) -- the documented large-donor input form -- passes the host-side checks of
    R/clone_id.R:84-96 (round 1 raised TypeError in the integer check) and only stops where the device is needed."""
    import scipy.sparse as sp
    import cardelino_b200 as cb
    rng = np.random.default_rng(0)
    D = sp.random(30, 40, density=0.3, random_state=1, data_rvs=lambda n: rng.integers(1, 9, n).astype(float)).tocsc()
    A = D.copy()
    A.data = np.floor(A.data * rng.random(A.nnz))
    counts = cb.SparseCounts.from_sparse(A, D)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        out = cb.clone_id(counts, None, n_clone=3, min_iter=5, max_iter=5, verbose=False)
        assert out["prob"].shape == (40, 3)
    else:
        with pytest.raises(cb.CdlError):
            cb.clone_id(counts, None, n_clone=3, min_iter=5, max_iter=5, verbose=False)


def test_committed_bench_lines_agree_on_the_parity_hashes():
    """The bench lines committed under profiles/ for 1, 2, 4 and 8 GPUs (C4, cells sharded) carry hashes of the labels
    of all 10^6 cells, the Config flips and the theta traces of the first sweeps of the bench chain: they must be the
    same at every GPU count, and every sharded line must say that it equals the unsharded donor rerun on rank 0."""
    import glob
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lines = {}
    for f in sorted(glob.glob(os.path.join(root, "profiles", "r2_bench_c4*.json"))):
        for ln in open(f):
            if ln.startswith("{"):
                d = json.loads(ln)
                if d.get("impl") != "reference" and d.get("parity") and d["config"]["workload"].startswith("C4:"):
                    lines[d["n_gpus"]] = d
    assert set(lines) >= {1, 2, 4, 8}, sorted(lines)
    one = lines[1]["parity"]
    for n, d in lines.items():
        p = d["parity"]
        assert p["sharded_equals_unsharded"] is True, n
        for key in ("labels_hash", "config_hash", "theta_hash"):
            assert p[key] == one[key], (n, key)
        if n > 1:
            assert p["label_mismatches"] == 0 and p["loglik_max_rel_diff"] < 1e-12
            for key in ("labels_hash", "config_hash", "theta_hash"):
                assert p["unsharded"][key] == one[key], (n, key)
