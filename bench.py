#!/usr/bin/env python
"""Benchmark of the clone-assignment Gibbs sweep (BASELINE.json metric: Gibbs iterations/sec on the
1M cells x 10k SNVs x 16 clones synthetic donor, cells sharded over N GPUs, strong scaling).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

One "step" = one full Gibbs sweep (R/clone_id.R:466-593) of one chain over the whole donor.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the byte accounting.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (N variants, M cells, K clones, coverage)
    "C4": (10000, 1000000, 16, 0.05),     # BASELINE.json configs[3], the metric's configuration
    "C3": (2000, 100000, 8, 0.19),        # configs[2] (sim_read_count-shaped 19 % coverage)
    "C4s8": (10000, 125000, 16, 0.05),    # one rank's share of C4 at 8 GPUs (tuning aid)
    "K32": (5000, 200000, 32, 0.05),      # widest clone count (separate mask table, 32 masked adds per entry)
    "N70k": (70000, 100000, 8, 0.01),     # more than 65536 variants: uint32 variant ids, weight table in global memory
    "tiny": (200, 4000, 4, 0.2),
}
SEED_DATA = 20261019 + 4
SEED_CHAIN = 1


def make_config(N, K, seed=SEED_DATA):
    """Random clonal tree Config (variants x clones, binary, base clone column all zero) like tree$Z."""
    rng = np.random.default_rng(seed)
    parent = [0] + [int(rng.integers(0, k)) for k in range(1, K)]
    origin = rng.integers(1, K, N) if K > 1 else np.zeros(N, int)     # clone in which the variant arises
    cfg = np.zeros((N, K))
    for k in range(1, K):
        anc = set()
        c = k
        while c != 0:
            anc.add(c)
            c = parent[c]
        cfg[:, k] = np.isin(origin, list(anc))
    return cfg


def workload_string(name):
    """config.workload: the same string in both arms (ours and --impl reference)."""
    if name == "C5":
        return ("C5: 64 independent chains on example_donor (34 variants x 428 cells x 4 clones), relax_Config=TRUE, "
                "min_iter=2000, max_iter=5000, Geweke stop rule, chain merge; plus one clone_id(inference='EM') call")
    N, M, K, cov = WORKLOADS[name]
    return ("%s: %d variants x %d cells x %d clones, %.0f%% coverage, wise=variant, relax_Config=TRUE (learned rate), "
            "1 chain" % (name, N, M, K, cov * 100))


def sha16(*arrays):
    import hashlib
    hh = hashlib.sha256()
    for a in arrays:
        hh.update(np.ascontiguousarray(a).tobytes())
    return hh.hexdigest()[:16]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[q] for r in self.rows if len(r) >= 6 for q in range(4) if r[2 + q] == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def ncu_traffic(workload, kernel):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of `kernel` from the committed
    `ncu --set full` capture of this workload, or None if there is no capture for it."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(workload, {}).get(kernel)
    except Exception:
        return None


def algorithmic_bytes(nnz, M, K, N):
    """Per sweep per chain: cell pass = nnz*(2+2+2) + 4(M+1) + M + 8MK (fp64 prob row);
    statistic pass = nnz*(2+2+2) + M."""
    s_idx = 2 if N <= 65536 else 4
    cells = nnz * (s_idx + 4) + 4 * (M + 1) + M + 8 * M * K
    stats = nnz * 6 + M
    return cells, stats


def run_ours(args):
    import torch
    import torch.distributed as dist
    import cardelino_b200 as cb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d" % args.gpus)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N, M, K, cov = WORKLOADS[args.workload]
    cfg = make_config(N, K)
    # contiguous cell shards (strong scaling: total work fixed)
    bounds = [(M * r) // world for r in range(world + 1)]
    off, Ml = bounds[rank], bounds[rank + 1] - bounds[rank]
    h = cb.Handle(local)
    if world > 1:
        import nvidia.nccl  # noqa: F401  (torch's bundled NCCL)
        nccl_path = os.path.join(list(nvidia.nccl.__path__)[0], "lib", "libnccl.so.2")
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.from_numpy(h.comm_unique_id(nccl_path)))
        dist.broadcast(uid, 0)
        h.comm_init(nccl_path, uid.cpu().numpy(), rank, world)
        h.set_exchange_mode(args.exchange)
    h.synth_generate(N, M, K, cfg, cov, SEED_DATA, cell_offset=off, M_local=Ml)
    _, _, nnz_local, _ = h.data_info()
    T0 = 4

    def prime(incremental):
        """A short run that leaves a chain on the device for the bench hook: -1 = the library default (incremental
        statistics wherever they apply), 0 = the full statistic pass every sweep."""
        h.gibbs_run(cfg, None, min_iter=T0, max_iter=T0, seed=SEED_CHAIN, keep_prob_all=1, wise="variant",
                    relax_Config=1, incremental_stats=incremental)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(incremental):
        """W untimed warm-up sweeps and the K timed sweeps are enqueued by ONE library call: the CUDA events that
        bracket the K sweeps are recorded on the library's stream right after the warm-up sweeps, whose per-sweep
        exchange has already aligned the ranks -- otherwise the few milliseconds of host-side skew between the
        ranks' Python processes would be charged to the first timed sweep.  barrier + synchronize on both sides,
        max over ranks."""
        prime(incremental)
        full0 = h.incremental_full_passes()
        barrier()
        r = h.bench_sweeps(args.warmup, args.steps, breakdown=False)
        barrier()
        ms = torch.tensor([r["ms_total"]], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        nb = max(3, min(args.steps, 10))
        bd = h.bench_sweeps(1, nb, breakdown=True)      # per-kernel durations (events around each kernel)
        return {"ms_total": float(ms.item()), "launches": int(r["launches"]),
                "full_passes": h.incremental_full_passes() - full0,
                "cells_ms": bd["ms_cells"] / nb, "stats_ms": bd["ms_stats"] / nb, "table_ms": bd["ms_table"] / nb}

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    main = timed(-1)            # the library's default path: this is `value`
    full = timed(0)             # every sweep re-reads all counts for the statistic table (round 1's `value`)
    nnz_t = torch.tensor([float(nnz_local)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(nnz_t, op=dist.ReduceOp.SUM)
    nnz = int(nnz_t.item())
    cells_b, stats_b = algorithmic_bytes(nnz_local, Ml, K, N)
    peak, peak_src = peaks()
    value = args.steps / (main["ms_total"] * 1e-3)

    # ---- parity: the sharded run against the unsharded donor on rank 0 (first sweeps, exact) ----
    parity = None
    if not args.no_parity:
        def first_sweeps(hh, mloc):
            hh.gibbs_run(cfg, None, min_iter=T0, max_iter=T0, seed=SEED_CHAIN, keep_prob_all=0, keep_assign_all=1,
                         wise="variant", relax_Config=1)
            z = hh.fetch("assign_all").astype(np.uint8)                  # T0 x M_local (row 1 is zeros)
            t0_, t1_ = hh.fetch("theta_traces")
            return z, hh.fetch("Config_all").astype(np.uint8), t0_, t1_, hh.fetch("logLik")
        z, c_all, th0, th1, ll = first_sweeps(h, Ml)
        if world > 1:
            zt = torch.zeros((T0, bounds[1] - bounds[0] + 1), dtype=torch.uint8, device="cuda")
            zt[:, :Ml] = torch.from_numpy(z).cuda()
            parts = [torch.empty_like(zt) for _ in range(world)] if rank == 0 else None
            dist.gather(zt, parts, dst=0)
            if rank == 0:
                z = np.concatenate([pp.cpu().numpy()[:, :bounds[r + 1] - bounds[r]] for r, pp in enumerate(parts)], axis=1)
        if rank == 0:
            parity = {"sweeps": T0 - 1, "labels_hash": sha16(z), "config_hash": sha16(c_all),
                      "theta_hash": sha16(th0, th1), "loglik_hash": sha16(ll),
                      "what": "sha256[:16] of assign_all (all %d cells in global order), Config_all and the theta0 / "
                              "theta1 traces of the first %d sweeps of the bench chain" % (M, T0 - 1)}
            if world > 1:
                h2 = cb.Handle(local)                  # the WHOLE donor on rank 0's GPU, no communicator
                h2.synth_generate(N, M, K, cfg, cov, SEED_DATA)
                z2, c2, a0, a1, l2 = first_sweeps(h2, M)
                h2.close()
                ref = {"labels_hash": sha16(z2), "config_hash": sha16(c2), "theta_hash": sha16(a0, a1)}
                parity["unsharded"] = dict(ref, loglik_hash=sha16(l2))
                # labels, Config flips and theta draws are exact (integer statistics, Philox keyed by global cell id); the
                # logLik trace adds each rank's sum(lchoose(D, A)) in rank order, so it agrees to rounding, not bitwise
                parity["sharded_equals_unsharded"] = all(parity[k] == ref[k] for k in ref)
                parity["label_mismatches"] = int((z != z2).sum())
                parity["loglik_max_rel_diff"] = float(np.max(np.abs(ll[1:] - l2[1:]) / np.abs(l2[1:])))
            else:
                parity["sharded_equals_unsharded"] = True
                parity["note"] = "one rank: the run IS the unsharded donor; compare the hashes across N in SCALE"
        barrier()
        prime(-1)

    # ---- end to end through the public API with host buffers (load + T sweeps + fetch) ----
    e2e = None
    if not args.no_e2e:
        def pinned(x):      # the donor sits in PINNED host memory, as the bench contract asks
            return torch.from_numpy(x.view(np.uint8)).pin_memory().numpy().view(x.dtype)

        p, i, a, d = [pinned(x) for x in h.export_csc_u16()]
        counts = cb.SparseCounts(N, Ml, p, i, a, d, cell_offset=off, M_total=M)
        T = args.e2e_iters
        # one short untimed call first: device allocations of this shape and the first-touch page faults of the host
        # result arrays are warm-up, like the W sweeps above
        cb.clone_id_Gibbs(counts, None, cfg, min_iter=20, max_iter=20, seed=SEED_CHAIN, verbose=False,
                          handle=h, keep_prob_all=0, light=True)
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        res = cb.clone_id_Gibbs(counts, None, cfg, min_iter=T, max_iter=T, seed=SEED_CHAIN, verbose=False,
                                handle=h, keep_prob_all=0, light=True)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        h2d = p.nbytes + i.nbytes + a.nbytes + d.nbytes + cfg.nbytes
        d2h = res["prob"].nbytes + res["Config_prob"].nbytes + res["theta1_all"].nbytes
        e2e = {"value": (T - 1) / float(dt.item()), "unit": "Gibbs iterations/s",
               "h2d_bytes_per_step": h2d / (T - 1), "d2h_bytes_per_step": d2h / (T - 1),
               "chain_iterations": T, "seconds": float(dt.item()),
               "note": "one clone_id_Gibbs call on pinned host uint16 CSC buffers: H2D + layout build + T sweeps + D2H "
                       "of prob / Config_prob / theta traces; T = --e2e-iters (default: the reference's default "
                       "min_iter); library defaults (incremental statistics on)"}
    clocks = sampler.stop() if sampler else None
    if clocks is not None:
        clocks["window"] = "timed sweeps (both modes), per-kernel passes, parity sweeps and the e2e call; 20 ms period"

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline_leg(args.workload, h.export_csc_u16())
        if not args.no_dense:
            cpu["dense_reference_shaped"] = cpu_dense_reference_shaped(args.workload, args.cpu_cells, args.cpu_iters)

    if rank == 0:
        kname = "k_cells_sell" if M >= 16384 else "k_cells"

        def roof(ms_launch, nbytes, kernel, traffic_key):
            return {"kernel": kernel, "achieved": nbytes / (ms_launch * 1e-3) / 1e9,
                    "frac": nbytes / (ms_launch * 1e-3) / 1e9 / peak,
                    "traffic": ncu_traffic(args.workload, traffic_key) if world == 1 else None,
                    "algorithmic_bytes_per_launch": nbytes, "ms_per_launch": ms_launch}
        r1 = roof(main["cells_ms"], cells_b, "%s (logLik + softmax + draw)" % kname, "k_cells_sell")
        out = {
            "metric": "Gibbs iterations/sec", "value": value, "unit": "Gibbs iterations/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": main["ms_total"] / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (device generator with sim_read_count's distributions)",
            "config": {"workload": workload_string(args.workload)},
            "nnz": nnz,
            "sharding": "cells over %d rank(s); per sweep the %d x %d statistic table is %s" %
                        (world, N, K, "not exchanged (1 rank)" if world == 1 else
                         ("summed over peer memory inside the table kernel (NVLink loads, no collective call)"
                          if h.exchange_is_peer() else "all-reduced with NCCL")),
            "l2": "inputs (%.1f GB per rank per pass) exceed the 126 MB L2" % (nnz_local * 6 / 1e9),
            "nnz_clone_updates_per_s": value * nnz * K,
            "statistics": {"mode": "incremental (library default): SA/SB corrected from the cells whose label changed "
                                   "when they are <= ~3 % of the cells, else the full pass; exact either way",
                           "full_statistic_passes": main["full_passes"],
                           "of_sweeps": args.steps + args.warmup + 1 + max(3, min(args.steps, 10))},
            "roofline": dict(r1, bound="hbm", peak=peak, unit="GB/s", peak_source=peak_src,
                             whole_sweep_frac=cells_b / (main["ms_total"] / args.steps * 1e-3) / 1e9 / peak,
                             statistics_ms=main["stats_ms"], table_kernel_ms=main["table_ms"]),
            "full_pass_mode": {
                "value": args.steps / (full["ms_total"] * 1e-3), "unit": "Gibbs iterations/s",
                "ms_per_step": full["ms_total"] / args.steps,
                "note": "incremental_stats = 0: every sweep reads the counts twice (cell pass + statistic pass); "
                        "round 1's `value`",
                "cells_kernel": roof(full["cells_ms"], cells_b, kname, "k_cells_sell"),
                "statistic_kernel": roof(full["stats_ms"], stats_b, "k_stats (per-variant x clone read sums)", "k_stats"),
                "table_kernel_ms": full["table_ms"]},
            "parity": parity, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
            "gpu_launches": main["launches"],
        }
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _cpu_dense_chain(job):
    """One reference-SHAPED chain (dense, op-for-op NumPy restatement of R/clone_id.R:466-593) on a cell sample;
    returns seconds per iteration (the first sweep, which pays the allocations, is not timed)."""
    workload, cells, iters, seed = job
    from oracle import cardelino_oracle as O
    N, M, K, cov = WORKLOADS[workload]
    rng = np.random.default_rng(seed)
    cfg = make_config(N, K)
    lab = rng.integers(0, K, cells)
    covered = rng.random((N, cells)) < cov
    D = np.where(covered, rng.choice([1, 1, 2, 2, 3, 4, 5, 8, 12, 30], size=(N, cells)), np.nan)
    th1 = rng.beta(0.45, 0.55, N)
    H = cfg[:, lab]
    p = np.where(H > 0, th1[:, None], 0.002)
    A = np.where(covered, rng.binomial(np.nan_to_num(D).astype(int), p), np.nan)
    times = []
    for T in (2, 2, iters + 2):       # rows 1..T: T - 1 sweeps.  The first run warms up (imports, page faults); the
        t0 = time.perf_counter()      # difference of the other two is `iters` sweeps without set-up and summaries
        O.clone_id_Gibbs(A, D, cfg, min_iter=T, max_iter=T, draws=O.NumpyDraws(seed))
        times.append(time.perf_counter() - t0)
    return max(times[2] - times[1], 1e-9) / iters


def cpu_dense_reference_shaped(workload, cells=None, iters=2):
    """The reference's own DENSE formulation (4K dense N x M double matrices rebuilt every sweep, R/clone_id.R:394-403,
    537-542) cannot hold C3 / C4 (51 GB / 5 TB of S-lists): it is timed on as many cells as fit a few GB -- 10 000 of
    C3's 100 000 cells, as BASELINE.md section 3 says, 1 000 of C4's 10^6 -- and scaled linearly in cells (every
    term of the dense sweep is O(N M K))."""
    N, M, K, cov = WORKLOADS[workload]
    if cells is None:
        cells = int(min(M, max(64, 6.4e8 // (4 * N * K))))      # S-lists of ~5 GB: 4K N x cells doubles
    sec = _cpu_dense_chain((workload, cells, iters, 100))
    return {"value": cells / M / sec, "unit": "Gibbs iterations/s", "cores": os.cpu_count(),
            "seconds_per_sweep_on_sample": sec,
            "sample": "%d sweeps of the dense R-shaped NumPy restatement (oracle/cardelino_oracle.py, BLAS threads as "
                      "NumPy finds them) on %d of %d cells, all %d variants and %d clones; scaled linearly in cells"
                      % (iters, cells, M, N, K)}


def cpu_sparse_port(workload, budget_s=20.0, counts=None, sweeps=None):
    """Best CPU figure: the sparse C / OpenMP restatement (oracle/gibbs_sparse.c, sufficient-statistic form, fp64)
    on ALL host threads (set explicitly: a launcher's OMP_NUM_THREADS=1 is ignored).  Runs whole sweeps on the
    full donor when they fit the budget, else on a cell sample (cost is linear in cells) -- says which."""
    from oracle import c_oracle as CO
    N, M, K, cov = WORKLOADS[workload]
    cfg_mask = CO.config_masks(make_config(N, K))
    threads = CO.set_threads()
    probe_cells = min(M, 20000)
    if counts is None:
        # no GPU in this process: make the donor on the host with the same distributions
        p, vi, a, d, _ = CO.synth(N, probe_cells, K, cov, cfg_mask, seed=SEED_DATA)
    else:
        p, vi, a, d = counts
        M = p.size - 1                     # this rank's cells
        probe_cells = min(M, probe_cells)
    sec, _, _ = CO.run(N, probe_cells, K, p[:probe_cells + 1], vi, a, d, cfg_mask, 2, seed=1)
    per_cell = sec / 2 / probe_cells
    want = sweeps if sweeps is not None else 3
    Ms = M if per_cell * M * want <= budget_s else max(probe_cells, int(budget_s / want / per_cell))
    if counts is None and Ms != probe_cells:
        p, vi, a, d, _ = CO.synth(N, Ms, K, cov, cfg_mask, seed=SEED_DATA)
    p = p[:Ms + 1]
    T = sweeps if sweeps is not None else max(2, min(20, int(budget_s / max(per_cell * Ms, 1e-9))))
    return {"N": N, "K": K, "M": M, "Ms": Ms, "T": T, "threads": threads, "run": lambda t: CO.run(N, Ms, K, p, vi, a, d, cfg_mask, t, seed=1)[0]}


def cpu_baseline_leg(workload, counts):
    c = cpu_sparse_port(workload, budget_s=20.0, counts=counts)
    sec = c["run"](c["T"])
    return {"value": c["T"] / sec * c["Ms"] / c["M"], "unit": "Gibbs iterations/s", "cores": c["threads"], "kind": "port",
            "sample": "%d sweeps of the sparse C/OpenMP restatement (oracle/gibbs_sparse.c, fp64) on %d of %d cells "
                      "(all %d variants, %d clones), %d threads%s" %
                      (c["T"], c["Ms"], c["M"], c["N"], c["K"], c["threads"],
                       "" if c["Ms"] == c["M"] else "; scaled linearly in cells to the full donor"),
            "seconds_per_sweep_on_sample": sec / c["T"]}


def run_reference(args):
    """The reference arm: R is not installed (here or on the GPU box) and the reference is pure R, so the reference's
    ALGORITHM is timed as its sparse C/OpenMP restatement (oracle/gibbs_sparse.c) on all host threads: W warm-up
    sweeps, then exactly K timed sweeps, each one full Gibbs sweep over the whole donor (a cell sample only if the
    host is too slow for the whole donor to finish in a few minutes; the line says which).  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "C5":
        return run_reference_c5(args)
    c = cpu_sparse_port(args.workload, budget_s=150.0, sweeps=args.steps + args.warmup)
    if args.warmup:
        c["run"](args.warmup)
    sec = c["run"](args.steps)
    value = args.steps / sec * c["Ms"] / c["M"]
    sample = ("%d sweeps of the sparse C/OpenMP restatement (oracle/gibbs_sparse.c, fp64) on %d of %d cells (all %d "
              "variants, %d clones), %d threads%s" % (args.steps, c["Ms"], c["M"], c["N"], c["K"], c["threads"],
                                                      "" if c["Ms"] == c["M"] else "; scaled linearly in cells"))
    out = {"impl": "reference", "metric": "Gibbs iterations/sec", "value": value, "unit": "Gibbs iterations/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / value,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic (host generator with sim_read_count's distributions)",
           "config": {"workload": workload_string(args.workload)},
           "reference_kind": "R is not installed: the reference ALGORITHM as its sparse C/OpenMP restatement on all "
                             "host threads (a far better CPU program than the reference's dense R loop; see "
                             "cpu_baseline.dense_reference_shaped in the other arm)",
           "cpu_baseline": {"value": value, "unit": "Gibbs iterations/s", "cores": c["threads"], "kind": "port",
                            "sample": sample},
           "e2e": {"value": value, "unit": "Gibbs iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


# ---- C5: 64 independent chains (one group per GPU) + the EM path + the Geweke stop rule ---------------------------
C5_CHAINS, C5_MIN, C5_MAX = 64, 2000, 5000


def _c5_donor():
    d = np.load(os.path.join(ROOT, "tests", "golden", "example_donor.npz"))
    return d["A"], d["D"], d["Z"]


def run_ours_c5(args):
    """BASELINE configs[4]: 64 chains of clone_id() on the reference's example_donor, one chain group per GPU
    (no communication while sampling), the reference's stop rule, chains merged like R/clone_id.R:165-181; plus
    one inference = "EM" call.  A step = one whole clone_id(n_chain = 64) call; value = chain iterations per second
    (the sum of the chains' n_iter over the call's wall time, merge included), max over ranks."""
    import torch
    import torch.distributed as dist
    import cardelino_b200 as cb
    from cardelino_b200 import distributed as cd
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % (29000 + os.getpid() % 2000), rank=0, world_size=1)
    A, D, Z = _c5_donor()
    h = cb.Handle(local)
    kw = dict(min_iter=C5_MIN, max_iter=C5_MAX, seed=SEED_CHAIN, handle=h, verbose=False)
    sampler = ClockSampler(local) if rank == 0 else None

    def one_call():
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        res = cd.clone_id_chain_groups(A, D, Z, C5_CHAINS, **kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda" if world > 1 else "cpu")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return res, float(t.item())
    for _ in range(max(1, min(args.warmup, 3))):
        one_call()
    if sampler:
        sampler.start()
    steps = max(1, min(args.steps, 20))
    tot_it, tot_s, res = 0, 0.0, None
    for _ in range(steps):
        res, dt = one_call()
        tot_it += sum(int(o["n_iter"]) - 1 for o in res["chains"])
        tot_s += dt
    clocks = sampler.stop() if sampler else None
    t0 = time.perf_counter()
    em = cb.clone_id(A, D, Config=Z, inference="EM", handle=h, verbose=False)
    em_s = time.perf_counter() - t0
    if rank == 0:
        value = tot_it / tot_s
        out = {"metric": "Gibbs iterations/sec", "value": value, "unit": "Gibbs iterations/s (summed over chains)",
               "n_gpus": world, "steps": steps, "warmup": max(1, min(args.warmup, 3)), "ms_per_step": tot_s / steps * 1e3,
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
               "data": "the reference's example_donor fixture (tests/golden/example_donor.npz)",
               "config": {"workload": workload_string("C5")},
               "chains": C5_CHAINS, "chain_iterations_per_call": tot_it / steps,
               "mean_n_iter": float(np.mean([o["n_iter"] for o in res["chains"]])),
               "em": {"seconds": em_s, "n_iter": int(em["n_iter"]), "logLik": float(em["logLik"])},
               "merged_prob_hash": sha16(np.round(res["prob"], 12)),
               "e2e": {"value": value, "unit": "Gibbs iterations/s (summed over chains)",
                       "h2d_bytes_per_step": int(A.nbytes + D.nbytes + Z.nbytes),
                       "d2h_bytes_per_step": int(sum(o["prob"].nbytes + o["Config_prob"].nbytes for o in res["chains"])),
                       "note": "value IS end to end: every step is a whole clone_id(n_chain = 64) call from host arrays "
                               "(dense load, 64 chains, per-chain fetches, host merge)"},
               "clocks": clocks, "gpu_launches": None, "roofline": None, "cpu_baseline": None}
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


def run_reference_c5(args):
    """C5 on the host: the NumPy restatement (op for op with the R loop) runs a few chains of the same call on one
    core each; chain iterations per second, scaled to the host's cores (chains are independent)."""
    from concurrent.futures import ProcessPoolExecutor
    cores = os.cpu_count() or 1
    n = min(cores, 8)
    t0 = time.perf_counter()
    with ProcessPoolExecutor(n) as ex:
        its = list(ex.map(_c5_cpu_chain, range(n)))
    dt = time.perf_counter() - t0
    value = sum(its) / dt * (cores / n)
    out = {"impl": "reference", "metric": "Gibbs iterations/sec", "value": value,
           "unit": "Gibbs iterations/s (summed over chains)", "n_gpus": args.gpus, "steps": 1, "warmup": 0,
           "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
           "data": "the reference's example_donor fixture", "config": {"workload": workload_string("C5")},
           "cpu_baseline": {"value": value, "unit": "Gibbs iterations/s", "cores": cores, "kind": "port",
                            "sample": "%d chains of 300 sweeps of the dense NumPy restatement, one process each; scaled "
                                      "to %d cores" % (n, cores)},
           "e2e": {"value": value, "unit": "Gibbs iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def _c5_cpu_chain(c):
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import cardelino_oracle as O
    A, D, Z = _c5_donor()
    O.clone_id_Gibbs(A, D, Z, min_iter=300, max_iter=300, draws=O.NumpyDraws(1000 + c))
    return 299


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C4", choices=list(WORKLOADS) + ["C5"])
    ap.add_argument("--e2e-iters", type=int, default=5000,
                    help="sweeps of the end-to-end clone_id_Gibbs call; 5000 = the reference's default min_iter "
                         "(R/clone_id.R:352), i.e. the shortest chain a stock clone_id() call runs")
    ap.add_argument("--cpu-cells", type=int, default=None,
                    help="cells of the dense R-shaped CPU sample (default: as many as give ~5 GB of S-lists)")
    ap.add_argument("--cpu-iters", type=int, default=1)
    ap.add_argument("--exchange", default="auto", choices=["auto", "nccl", "peer"],
                    help="multi-GPU statistic-table exchange: peer memory fused into the table kernel, or NCCL")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-dense", action="store_true", help="skip the dense R-shaped CPU sample (about a minute)")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    # stdout carries exactly ONE JSON line: anything libraries print while we work (NCCL's version banner,
    # torchrun notices) goes to stderr instead
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    lines = []
    real_print = print

    def capture(*a, **k):
        if k.get("file") in (None, sys.stdout):
            lines.append(" ".join(str(x) for x in a))
        else:
            real_print(*a, **k)

    import builtins
    builtins.print = capture
    try:
        if args.impl == "reference":
            run_reference(args)
        elif args.workload == "C5":
            run_ours_c5(args)
        else:
            run_ours(args)
    finally:
        builtins.print = real_print
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    for ln in lines:
        if ln.startswith("{"):
            real_print(ln, flush=True)
        else:
            real_print(ln, file=sys.stderr, flush=True)


if __name__ == "__main__":
    main()
