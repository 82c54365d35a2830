"""Multi-GPU partitioning of the lattice path.  Scenarios (config 4) and the
candidates of one dense query (config 5) are independent, so ranks take contiguous blocks and
the only exchange is the final gather of one (cost, index) pair per rank -- 8 bytes."""
import numpy as np


def block(n, rank, world):
    """contiguous block [lo, hi) of n items for `rank` of `world` (sizes differ by at most 1)"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_best(cost, idx, group=None):
    """lexicographic (cost, idx) minimum over ranks == np.argmin's first-minimum rule on the
    concatenated cost vector.  Works on any torch.distributed backend (gloo on CPU, nccl)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(cost), int(idx)
    world = dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" \
        else torch.device("cpu")
    mine = torch.tensor([float(cost), float(idx)], dtype=torch.float64, device=dev)
    allp = [torch.zeros(2, dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(allp, mine, group=group)
    pairs = [(float(p[0]), int(p[1])) for p in allp]
    return min(pairs)


def gather_stats(values, group=None):
    """sum / max over ranks of a small float vector (throughput statistics)"""
    import torch
    import torch.distributed as dist
    v = np.asarray(values, dtype=np.float64)
    if not (dist.is_available() and dist.is_initialized()):
        return v, v
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" \
        else torch.device("cpu")
    s = torch.tensor(v, device=dev)
    m = s.clone()
    dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    return s.cpu().numpy(), m.cpu().numpy()
