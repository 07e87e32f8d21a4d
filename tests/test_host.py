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


def test_r_shim_compiles_against_stub_headers():
    """The .Call shim cannot be built here (no R), but it must at least be valid C against the C ABI header: compile
    it with -fsyntax-only -Wall -Werror against stand-in declarations of the R API entry points it uses."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(root, "tests", "r_stub"),
                        "-I", os.path.join(root, "include"), os.path.join(root, "r", "cardelino_b200_shim.c")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
