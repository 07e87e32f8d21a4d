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
                                          "--format=csv,noheader,nounits", "-lms", "100"],
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
    h.gibbs_run(cfg, None, min_iter=T0, max_iter=T0, seed=SEED_CHAIN, keep_prob_all=1, wise="variant",
                relax_Config=1)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    # W untimed warm-up sweeps and the K timed sweeps are enqueued by ONE library call: the CUDA events that
    # bracket the K sweeps are recorded on the stream right after the warm-up sweeps, whose per-sweep exchange
    # has already aligned the ranks -- otherwise the few milliseconds of host-side skew between the ranks'
    # Python processes would be charged to the first timed sweep.  barrier + synchronize on both sides.
    r = h.bench_sweeps(args.warmup, args.steps, breakdown=False)
    barrier()
    clocks = sampler.stop() if sampler else None
    ms = torch.tensor([r["ms_total"]], dtype=torch.float64, device="cuda")
    nnz_t = torch.tensor([float(nnz_local)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(nnz_t, op=dist.ReduceOp.SUM)
    ms_total = float(ms.item())
    nnz = int(nnz_t.item())
    # per-kernel durations (second pass; events around each kernel on the library's stream)
    bd = h.bench_sweeps(1, max(3, min(args.steps, 10)), breakdown=True)
    nb = max(3, min(args.steps, 10))
    cells_b, stats_b = algorithmic_bytes(nnz_local, Ml, K, N)
    peak, peak_src = peaks()
    cells_ms = bd["ms_cells"] / nb
    stats_ms = bd["ms_stats"] / nb
    value = args.steps / (ms_total * 1e-3)

    # ---- optional mode: incremental statistics (exact; reported beside, never instead of, `value`) ----
    incr = None
    if world == 1 and not args.no_incremental:
        h.gibbs_run(cfg, None, min_iter=T0, max_iter=T0, seed=SEED_CHAIN, keep_prob_all=1, wise="variant",
                    relax_Config=1, incremental_stats=1)
        full0 = h.incremental_full_passes()
        ri = h.bench_sweeps(args.warmup, args.steps, breakdown=False)
        incr = {"value": args.steps / (ri["ms_total"] * 1e-3), "unit": "Gibbs iterations/s",
                "ms_per_step": ri["ms_total"] / args.steps,
                "full_statistic_passes": h.incremental_full_passes() - full0,
                "of_sweeps": args.steps + args.warmup,
                "note": "cdl_gibbs_params.incremental_stats = 1: SA/SB corrected from the cells whose label changed "
                        "(integer sums, bit-identical chain) when they are <= ~3 % of the cells; on this synthetic "
                        "donor every cell has ~500 covered variants, so labels are settled after the first sweeps"}
        h.gibbs_run(cfg, None, min_iter=T0, max_iter=T0, seed=SEED_CHAIN, keep_prob_all=1, wise="variant",
                    relax_Config=1)

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
               "note": "one clone_id_Gibbs call on host uint16 CSC buffers: H2D + layout build + T sweeps + D2H; "
                       "T = --e2e-iters (default: the reference's default min_iter)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_sparse_port(args.workload, budget_s=20.0, counts=h.export_csc_u16())
        cpu["dense_reference_shaped"] = cpu_dense_reference_shaped(args.workload, args.cpu_cells, args.cpu_iters)

    if rank == 0:
        out = {
            "metric": "Gibbs iterations/sec", "value": value, "unit": "Gibbs iterations/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (device generator with sim_read_count's distributions)",
            "config": {"workload": "%s: %d variants x %d cells x %d clones, %.0f%% coverage, nnz=%d, "
                                   "wise=variant, relax_Config=TRUE (learned rate), 1 chain" %
                                   (args.workload, N, M, K, cov * 100, nnz),
                       "sharding": "cells over %d rank(s); per sweep the %d x %d statistic table is %s" %
                                   (world, N, K, "not exchanged (1 rank)" if world == 1 else
                                    ("summed over peer memory inside the table kernel (NVLink loads, no collective call)"
                                     if h.exchange_is_peer() else "all-reduced with NCCL")),
                       "l2": "inputs (%.1f GB per rank per pass) exceed the 126 MB L2" % (nnz_local * 6 / 1e9)},
            "nnz_clone_updates_per_s": value * nnz * K,
            "roofline": {"bound": "hbm", "kernel": "%s (logLik + softmax + draw)" % ("k_cells_sell" if M >= 16384 else "k_cells"),
                         "achieved": cells_b / (cells_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": cells_b / (cells_ms * 1e-3) / 1e9 / peak,
                         "traffic": ncu_traffic(args.workload, "k_cells_sell") if world == 1 else None,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": cells_b,
                         "ms_per_launch": cells_ms,
                         "second_kernel": {"kernel": "k_stats (per-variant x clone read sums)",
                                           "achieved": stats_b / (stats_ms * 1e-3) / 1e9,
                                           "frac": stats_b / (stats_ms * 1e-3) / 1e9 / peak,
                                           "traffic": ncu_traffic(args.workload, "k_stats") if world == 1 else None,
                                           "algorithmic_bytes_per_launch": stats_b, "ms_per_launch": stats_ms},
                         "table_kernel_ms": bd["ms_table"] / nb},
            "cpu_baseline": cpu, "e2e": e2e, "incremental_stats_mode": incr, "clocks": clocks,
            "gpu_launches": int(r["launches"]),
        }
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _cpu_dense_chain(job):
    """One reference-SHAPED chain (dense, op-for-op NumPy restatement of R/clone_id.R:466-593) on a cell sample;
    returns seconds per iteration."""
    workload, cells, iters, seed = job
    os.environ.setdefault("OMP_NUM_THREADS", "1")
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
    T = iters + 2
    t0 = time.perf_counter()
    O.clone_id_Gibbs(A, D, cfg, min_iter=T, max_iter=T, draws=O.NumpyDraws(seed))
    dt = time.perf_counter() - t0
    return dt / (T - 1)


def cpu_dense_reference_shaped(workload, cells=64, iters=2):
    """The reference's own DENSE formulation (4K dense N x M matrices, R/clone_id.R:394-403) cannot hold C3 / C4
    (SURVEY.md fact 4); timed on a few cells, cost is linear in cells."""
    N, M, K, cov = WORKLOADS[workload]
    sec = _cpu_dense_chain((workload, cells, iters, 100))
    return {"value": cells / M / sec, "unit": "Gibbs iterations/s", "cores": 1,
            "sample": "%d sweeps of the dense R-shaped NumPy restatement on %d of %d cells, scaled linearly in "
                      "cells" % (iters, cells, M)}


def cpu_sparse_port(workload, budget_s=20.0, counts=None):
    """Best CPU figure: the sparse C / OpenMP restatement (oracle/gibbs_sparse.c, sufficient-statistic form, fp64)
    on all host threads.  Runs whole sweeps on the full donor when one sweep fits the budget, else on a cell
    sample (cost is linear in cells) -- says which."""
    from oracle import c_oracle as CO
    N, M, K, cov = WORKLOADS[workload]
    cfg_mask = CO.config_masks(make_config(N, K))
    threads = CO.threads()
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
    Ms = M if per_cell * M * 3 <= budget_s else max(probe_cells, int(budget_s / 3 / per_cell))
    if counts is None and Ms != probe_cells:
        p, vi, a, d, _ = CO.synth(N, Ms, K, cov, cfg_mask, seed=SEED_DATA)
    p = p[:Ms + 1]
    T = max(2, min(20, int(budget_s / max(per_cell * Ms, 1e-9))))
    sec, _, _ = CO.run(N, Ms, K, p, vi, a, d, cfg_mask, T, seed=1)
    return {"value": T / sec * Ms / M, "unit": "Gibbs iterations/s", "cores": threads, "kind": "port",
            "sample": "%d sweeps of the sparse C/OpenMP restatement (oracle/gibbs_sparse.c, fp64) on %d of %d cells "
                      "(all %d variants, %d clones), %d threads%s" %
                      (T, Ms, M, N, K, threads, "" if Ms == M else "; scaled linearly in cells to the full donor"),
            "seconds_per_sweep_on_sample": sec / T}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    N, M, K, cov = WORKLOADS[args.workload]
    t0 = time.perf_counter()
    vals, cpu = [], None
    budget = max(5.0, 120.0 / max(1, args.steps + args.warmup))
    for s in range(args.warmup + args.steps):
        c = cpu_sparse_port(args.workload, budget_s=budget)
        if s >= args.warmup:
            vals.append(c["value"])
            cpu = c
        if time.perf_counter() - t0 > 150:
            break
    if not vals:
        cpu = cpu_sparse_port(args.workload, budget_s=5.0)
        vals = [cpu["value"]]
    value = float(np.mean(vals))
    cpu = dict(cpu, value=value)
    dense = cpu_dense_reference_shaped(args.workload)
    out = {"impl": "reference", "metric": "Gibbs iterations/sec", "value": value, "unit": "Gibbs iterations/s",
           "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup,
           "ms_per_step": 1e3 / value if value else None, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic (host generator with sim_read_count's distributions)",
           "config": {"workload": "%s: %d variants x %d cells x %d clones, %.0f%% coverage, wise=variant, "
                                  "relax_Config=TRUE, 1 chain.  R is not installed here: the reference ALGORITHM is "
                                  "timed as the sparse C/OpenMP restatement on all host threads (the best CPU figure); "
                                  "its dense R-shaped formulation is in dense_reference_shaped" %
                                  (args.workload, N, M, K, cov * 100)},
           "cpu_baseline": cpu, "dense_reference_shaped": dense,
           "e2e": {"value": value, "unit": "Gibbs iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C4", choices=list(WORKLOADS))
    ap.add_argument("--e2e-iters", type=int, default=5000,
                    help="sweeps of the end-to-end clone_id_Gibbs call; 5000 = the reference's default min_iter "
                         "(R/clone_id.R:352), i.e. the shortest chain a stock clone_id() call runs")
    ap.add_argument("--cpu-cells", type=int, default=64)
    ap.add_argument("--cpu-iters", type=int, default=2)
    ap.add_argument("--exchange", default="auto", choices=["auto", "nccl", "peer"],
                    help="multi-GPU statistic-table exchange: peer memory fused into the table kernel, or NCCL")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-incremental", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
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
