"""world_size-2 gloo test of the multi-rank host logic (cell sharding): shards partition the cells, the
per-shard read-sum tables add up to the global table (the quantity the library all-reduces with NCCL
every sweep), Philox draws keyed by GLOBAL cell id are identical however the cells are split, and the
per-shard results gather back into the full matrices."""
import os

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    import sys
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from cardelino_b200 import distributed as D
    from oracle import cardelino_oracle as O
    from oracle import philox
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = np.load(os.path.join(ROOT, "tests", "golden", "example_donor.npz"))
    A, Dm, Z = d["A"], d["D"], d["Z"]
    A1, B1, W = O.preprocess(A, Dm)
    M, K = A.shape[1], Z.shape[1]
    bounds = D.shard_bounds(M, world)
    lo, hi = bounds[rank]
    # one sweep, sharded: logLik/softmax/draw are per cell; only SA/SB need the exchange
    theta0, theta1 = 0.02, np.linspace(0.3, 0.7, A.shape[0])
    u = philox.uniforms(5, 0, 2, philox.STREAM_ASSIGN, np.arange(lo, hi))         # global cell ids
    prob, lab = O.sweep_tables(A1[:, lo:hi], B1[:, lo:hi], Z, theta0, theta1, np.full(K, 1 / K), u)
    Zl = np.zeros((hi - lo, K))
    Zl[np.arange(hi - lo), lab - 1] = 1
    sab = np.stack([A1[:, lo:hi] @ Zl, B1[:, lo:hi] @ Zl])
    sab = D.allreduce_sum(sab)
    prob_full = D.gather_rows(prob, bounds)
    wsum = D.allreduce_sum(np.array([O.preprocess(A[:, lo:hi], Dm[:, lo:hi])[2]]))[0]
    uid = D.broadcast_bytes(np.arange(128, dtype=np.uint8) if rank == 0 else np.zeros(128, np.uint8))
    if rank == 0:
        np.savez(out, sab=sab, prob=prob_full, wsum=wsum, uid=uid)
    dist.barrier()
    dist.destroy_process_group()


def test_cell_sharding_world2(tmp_path):
    from cardelino_b200 import distributed as D
    from oracle import cardelino_oracle as O
    from oracle import philox
    assert D.shard_bounds(10, 3) == [(0, 3), (3, 6), (6, 10)]
    b = D.shard_bounds(1000003, 8)
    assert b[0][0] == 0 and b[-1][1] == 1000003 and all(b[i][1] == b[i + 1][0] for i in range(7))
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(2, 29511, out), nprocs=2, join=True)
    r = np.load(out)
    d = np.load(os.path.join(ROOT, "tests", "golden", "example_donor.npz"))
    A, Dm, Z = d["A"], d["D"], d["Z"]
    A1, B1, W = O.preprocess(A, Dm)
    M, K = A.shape[1], Z.shape[1]
    u = philox.uniforms(5, 0, 2, philox.STREAM_ASSIGN, np.arange(M))
    prob, lab = O.sweep_tables(A1, B1, Z, 0.02, np.linspace(0.3, 0.7, 34), np.full(K, 1 / K), u)
    Zl = np.zeros((M, K))
    Zl[np.arange(M), lab - 1] = 1
    assert np.array_equal(r["sab"], np.stack([A1 @ Zl, B1 @ Zl]))     # integer sums: exact at any rank count
    assert np.allclose(r["prob"], prob, rtol=0, atol=1e-13)   # BLAS blocking differs with the shard shape
    assert abs(r["wsum"] - W) < 1e-9
    assert np.array_equal(r["uid"], np.arange(128, dtype=np.uint8))


def _chain_worker(rank, world, port, out):
    import sys
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from cardelino_b200 import distributed as D
    from oracle import cardelino_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = np.load(os.path.join(ROOT, "tests", "golden", "example_donor.npz"))
    A, Dm, Z = d["A"], d["D"], d["Z"]

    def run_chains(first, count):      # the oracle stands in for the GPU library on this CPU-only box
        return [O.clone_id_Gibbs(A, Dm, Z, min_iter=60, max_iter=60, relax_Config=False, draws=O.NumpyDraws(500 + c))
                for c in range(first, first + count)]

    res = D.clone_id_chain_groups(A, Dm, Z, n_chain=3, run_chains=run_chains)
    if rank == 0:
        np.savez(out, prob=res["prob"], n_chain=res["n_chain"], relax=res["relax_rate"], cfg=res["Config_prob"])
    dist.barrier()
    dist.destroy_process_group()


def test_chain_groups_world2(tmp_path):
    """Chains dealt to two ranks (2 + 1), gathered and merged like R/clone_id.R:165-181: the same result as merging
    the three chains in one process."""
    from cardelino_b200 import distributed as D
    import importlib
    CI = importlib.import_module("cardelino_b200.clone_id")
    from oracle import cardelino_oracle as O
    assert D.chain_groups(64, 8) == [(8 * r, 8) for r in range(8)] and D.chain_groups(3, 2) == [(0, 1), (1, 2)]
    out = str(tmp_path / "chains.npz")
    mp.spawn(_chain_worker, args=(2, 29517, out), nprocs=2, join=True)
    r = np.load(out)
    d = np.load(os.path.join(ROOT, "tests", "golden", "example_donor.npz"))
    A, Dm, Z = d["A"], d["D"], d["Z"]
    outs = [O.clone_id_Gibbs(A, Dm, Z, min_iter=60, max_iter=60, relax_Config=False, draws=O.NumpyDraws(500 + c))
            for c in range(3)]
    ref = CI.merge_chains(outs)
    assert int(r["n_chain"]) == 3 and np.allclose(r["prob"], ref["prob"], atol=1e-15)
    assert np.allclose(r["cfg"], ref["Config_prob"]) and abs(float(r["relax"]) - ref["relax_rate"]) < 1e-15
