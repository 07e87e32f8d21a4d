"""Host plumbing for cell-sharded multi-GPU runs (one process per GPU, torch.distributed).

The data path has exactly one exchange per sweep -- an all-reduce of the N x K read-sum table -- and it
runs inside the library on NCCL (cdl_comm_init).  torch.distributed is only used here to hand every rank
the ncclUniqueId and to gather the per-shard results; with the gloo backend the same code runs on CPU
(tests/test_distributed_cpu.py).
"""
import os

import numpy as np


def shard_bounds(M, world):
    """Contiguous cell ranges [lo, hi) per rank; sizes differ by at most one cell."""
    edges = [(M * r) // world for r in range(world + 1)]
    return [(edges[r], edges[r + 1]) for r in range(world)]


def nccl_library_path():
    """torch's bundled libnccl (the copy torch.distributed already loaded in this process)."""
    try:
        import nvidia.nccl
        p = os.path.join(list(nvidia.nccl.__path__)[0], "lib", "libnccl.so.2")
        if os.path.exists(p):
            return p
    except Exception:
        pass
    return "libnccl.so.2"


def broadcast_bytes(buf, src=0):
    """Broadcast a uint8 numpy buffer from `src` with whatever backend the default group uses."""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.from_numpy(np.ascontiguousarray(buf, dtype=np.uint8)).to(dev)
    dist.broadcast(t, src)
    return t.cpu().numpy()


def init_comm(handle, rank=None, world=None):
    """Create the library's NCCL communicator on every rank (unique id made on rank 0)."""
    import torch.distributed as dist
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    path = nccl_library_path()
    uid = handle.comm_unique_id(path) if rank == 0 else np.zeros(128, np.uint8)
    uid = broadcast_bytes(uid, 0)
    handle.comm_init(path, uid, rank, world)
    return path


def gather_rows(local, bounds):
    """All-gather per-shard row blocks (cells x K) into the full matrix on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    K = local.shape[1]
    mx = max(hi - lo for lo, hi in bounds)
    pad = np.zeros((mx, K))
    pad[: local.shape[0]] = local
    t = torch.from_numpy(pad).to(dev)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    return np.concatenate([o.cpu().numpy()[: hi - lo] for o, (lo, hi) in zip(outs, bounds)], axis=0)


def allreduce_sum(arr):
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.from_numpy(np.ascontiguousarray(arr)).to(dev)
    dist.all_reduce(t)
    return t.cpu().numpy()


def chain_groups(n_chain, world):
    """Chains dealt to ranks in contiguous groups: [(first chain, count)] per rank (counts differ by at most 1)."""
    edges = [(n_chain * r) // world for r in range(world + 1)]
    return [(edges[r], edges[r + 1] - edges[r]) for r in range(world)]


def clone_id_chain_groups(A, D, Config, n_chain, run_chains=None, **kwargs):
    """Independent chains placed one group per GPU (BASELINE config 5; the reference forks one process per chain,
    R/clone_id.R:127-154): every rank holds the whole donor, runs its group of chains with globally unique chain
    numbers (Philox keys), the per-chain prob / Config_prob / relax_rate are gathered and merged exactly like
    clone_id() merges them (R/clone_id.R:165-181).  No communication while sampling.  Returns the merged result on
    every rank.  `run_chains(first, count)` -> list of per-chain dicts replaces the GPU call in tests."""
    import torch.distributed as dist
    import importlib
    _ci = importlib.import_module(__package__ + ".clone_id")     # the module (the package re-exports the function)
    rank, world = dist.get_rank(), dist.get_world_size()
    first, count = chain_groups(n_chain, world)[rank]
    kwargs.setdefault("light", True)       # only prob / Config_prob / scalars are gathered: skip the big traces
    if run_chains is None:
        def run_chains(first, count):
            if count == 0:
                return []
            return _ci.clone_id_Gibbs(A, D, Config, n_chain=count, chain_offset=first, _all_chains=True, **kwargs)
    mine = run_chains(first, count)
    keep = ("prob", "Config_prob", "relax_rate", "theta0", "n_iter", "percent_converged", "logLik_post")
    slim = [{k: o[k] for k in keep if k in o} for o in mine]
    gathered = [None] * world
    dist.all_gather_object(gathered, slim)
    outs = [o for part in gathered for o in part]
    merged = _ci.merge_chains(outs)
    merged["chains"] = outs
    return merged
