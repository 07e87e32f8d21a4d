"""Worker of tests/test_multi_gpu.py (run under torch.distributed.run, one rank per GPU).

Every rank holds a contiguous cell shard of one synthetic donor.  The sharded chain is run twice -- with
the NCCL all-reduce of the statistic table and with the peer-memory exchange fused into the table kernel --
and rank 0 also runs the WHOLE donor on its own GPU without any communicator.  All three must give the
same labels, Config flips, theta traces and logLik trace bit for bit (integer statistics, Philox keyed by
global cell id)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import cardelino_b200 as cb
    from cardelino_b200 import distributed as D
    import bench

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N, M, K, cov, T = 600, 50000, 8, 0.1, 10
    cfg = bench.make_config(N, K, seed=3)
    lo, hi = D.shard_bounds(M, world)[rank]
    h = cb.Handle(local)
    D.init_comm(h)
    h.synth_generate(N, M, K, cfg, cov, 99, cell_offset=lo, M_local=hi - lo)
    out = {}
    full_passes = {}
    # nccl: all-reduce per sweep (full statistic pass); peer: tables summed over peer memory inside the table kernel,
    # incremental statistics on (the library default); peer_full: the same exchange with the full pass every sweep
    for mode in ("nccl", "peer", "peer_full"):
        h.set_exchange_mode(mode.split("_")[0])
        h.gibbs_run(cfg, None, min_iter=T, max_iter=T, seed=4, keep_prob_all=1, keep_assign_all=1,
                    incremental_stats=0 if mode == "peer_full" else -1)
        assert h.exchange_is_peer() == (mode != "nccl"), mode
        full_passes[mode] = h.incremental_full_passes()
        z = torch.from_numpy(h.fetch("assign_all").astype(np.int32)).cuda()          # T x M_local
        parts = [torch.empty((T, b - a), dtype=torch.int32, device="cuda") for a, b in D.shard_bounds(M, world)]
        dist.all_gather(parts, z)
        out[mode] = {"z": torch.cat(parts, dim=1).cpu().numpy(), "cfg": h.fetch("Config_all"),
                     "ll": h.fetch("logLik"), "t0": h.fetch("theta_traces")[0], "t1": h.fetch("theta_traces")[1]}
    ok = True
    msg = []
    if rank == 0:
        full = cb.Handle(local)
        full.synth_generate(N, M, K, cfg, cov, 99)
        full.gibbs_run(cfg, None, min_iter=T, max_iter=T, seed=4, keep_prob_all=1, keep_assign_all=1)
        ref = {"z": full.fetch("assign_all"), "cfg": full.fetch("Config_all"), "ll": full.fetch("logLik"),
               "t0": full.fetch("theta_traces")[0], "t1": full.fetch("theta_traces")[1]}
        for mode in out:
            for key in ("z", "cfg", "t0", "t1"):
                if not np.array_equal(out[mode][key], ref[key]):
                    ok = False
                    msg.append("%s/%s differs from the single-GPU chain" % (mode, key))
            # logLik: each rank adds its shard's W_log in a different order -> equal to rounding
            if not np.allclose(out[mode]["ll"], ref["ll"], rtol=1e-12, atol=0):
                ok = False
                msg.append("%s/logLik differs" % mode)
        if not (1 <= full_passes["peer"] <= T - 1) or full_passes["nccl"] != 0 or full_passes["peer_full"] != 0:
            ok = False
            msg.append("incremental statistics did not run as expected: %s" % full_passes)
        print(json.dumps({"ok": ok, "world": world, "messages": msg, "full_statistic_passes": full_passes}))
    # ---- the stop rule on a sharded chain: the Geweke counts are summed over the ranks, so every rank stops at the
    # iteration the unsharded chain stops at -- with the whole trace (exact mode) and from the running sums and
    # window snapshots of streaming mode ----
    stop = {}
    for keep in (1, 0):
        h.gibbs_run(cfg, None, min_iter=100, max_iter=600, seed=8, keep_prob_all=keep)
        info = h.gibbs_info(0)
        stop[keep] = (int(info.n_iter), int(info.n_buin), float(info.percent_converged), h.fetch("prob"))
    if rank == 0:
        for keep in (1, 0):
            full.gibbs_run(cfg, None, min_iter=100, max_iter=600, seed=8, keep_prob_all=keep)
            fi = full.gibbs_info(0)
            n_it, n_bu, pc, prob = stop[keep]
            if (n_it, n_bu) != (int(fi.n_iter), int(fi.n_buin)) or abs(pc - float(fi.percent_converged)) > 0.2:
                ok = False
                msg.append("stop rule (keep_prob_all=%d): sharded %s vs unsharded %s" %
                           (keep, (n_it, n_bu, pc), (int(fi.n_iter), int(fi.n_buin), float(fi.percent_converged))))
            elif not np.allclose(prob, full.fetch("prob")[lo:hi], rtol=0, atol=1e-9):
                ok = False
                msg.append("posterior means differ (keep_prob_all=%d)" % keep)
        if stop[1][:2] != stop[0][:2]:
            ok = False
            msg.append("streaming and exact mode stop at different iterations: %s vs %s" % (stop[0][:2], stop[1][:2]))
        print(json.dumps({"ok": ok, "world": world, "messages": msg, "stop": [stop[1][:3], stop[0][:3]]}))
    # ---- wise = "element" on cell shards (R/clone_id.R:413-416,557-571): every rank owns the theta1 of its own entries, the
    # two N x K tables are all-reduced, the Philox index of an entry is its position in the whole donor ----
    Ne, Me, Ke, Te = 120, 6000, 4, 12
    cfg_e = bench.make_config(Ne, Ke, seed=5)
    lo_e, hi_e = D.shard_bounds(Me, world)[rank]
    he = cb.Handle(local)
    D.init_comm(he)
    he.synth_generate(Ne, Me, Ke, cfg_e, 0.3, 77, cell_offset=lo_e, M_local=hi_e - lo_e)
    he.gibbs_run(cfg_e, None, min_iter=Te, max_iter=Te, seed=9, keep_prob_all=1, keep_assign_all=1, wise="element")
    mine = {"z": he.fetch("assign_all"), "cfg": he.fetch("Config_all"), "t0": he.fetch("theta_traces")[0],
            "t1": he.fetch("theta_traces")[1], "prob": he.fetch("prob"), "ll": he.fetch("logLik")}
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        fe = cb.Handle(local)
        fe.synth_generate(Ne, Me, Ke, cfg_e, 0.3, 77)
        fe.gibbs_run(cfg_e, None, min_iter=Te, max_iter=Te, seed=9, keep_prob_all=1, keep_assign_all=1, wise="element")
        t0r, t1r = fe.fetch("theta_traces")
        checks = {
            "labels": np.array_equal(np.concatenate([g["z"] for g in gathered], axis=1), fe.fetch("assign_all")),
            "Config_all": all(np.array_equal(g["cfg"], fe.fetch("Config_all")) for g in gathered),
            "theta0": all(np.array_equal(g["t0"], t0r) for g in gathered),
            "theta1 (per entry, shards side by side)": np.array_equal(np.concatenate([g["t1"] for g in gathered], axis=1), t1r),
            "prob": np.allclose(np.concatenate([g["prob"] for g in gathered], axis=0), fe.fetch("prob"), rtol=0, atol=1e-9),
            "logLik": all(np.allclose(g["ll"], fe.fetch("logLik"), rtol=1e-10, atol=0) for g in gathered),
        }
        for name, good in checks.items():
            if not good:
                ok = False
                msg.append("wise=element: sharded %s differs from the single-GPU chain" % name)
        print(json.dumps({"ok": ok, "world": world, "messages": msg, "element_entries": int(t1r.shape[1])}))
        fe.close()
    he.close()
    # ---- independent chains, one group per GPU (BASELINE config 5): the merged result equals the one of the same
    # chains (same Philox chain numbers) run in a single process ----
    d = np.load(os.path.join(ROOT, "tests", "golden", "example_donor.npz"))
    A, Dm, Z = d["A"], d["D"], d["Z"]
    hc = cb.Handle(local)
    kw = dict(min_iter=80, max_iter=80, seed=6, relax_Config=False, verbose=False, keep_prob_all=1)
    res = D.clone_id_chain_groups(A, Dm, Z, n_chain=4, handle=hc, **kw)
    if rank == 0:
        one = cb.clone_id_Gibbs(A, Dm, Z, n_chain=4, handle=hc, _all_chains=True, **kw)
        import importlib
        ref = importlib.import_module("cardelino_b200.clone_id").merge_chains(one)
        same = res["n_chain"] == 4 and np.array_equal(res["prob"], ref["prob"]) and \
            all(np.array_equal(a["prob"], b["prob"]) for a, b in zip(res["chains"], one))
        if not same:
            ok = False
            msg.append("chain groups differ from the single-process chains")
        print(json.dumps({"ok": ok, "world": world, "messages": msg}))
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
