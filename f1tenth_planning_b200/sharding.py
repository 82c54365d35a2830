"""Multi-GPU partitioning of the lattice path.  Scenarios (config 4) and the
candidates of one dense query (config 5) are independent, so ranks take contiguous blocks and
the only exchange is the final gather of one (cost, index) pair per rank -- 8 bytes.

Two ways to do that exchange for a candidate-sharded query:
  * Engine.attach_peers(): the select kernel itself pushes the rank's packed (cost, index) key
    into every peer's HBM over NVLink (CUDA IPC mapped peer memory, system-scope atomicMin) and
    every rank's plan(shard=...) returns the global winner -- no host round trip, no collective
    library call on the latency path (GPU ranks of one node);
  * reduce_best(): a torch.distributed all-gather of the pair plus a host min -- any backend
    (gloo on CPU), used by the CPU tests and as the cross-check of the first."""
import numpy as np


def block(n, rank, world):
    """contiguous block [lo, hi) of n items for `rank` of `world` (sizes differ by at most 1)"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class _Workspace:
    """preallocated buffers of reduce_best (one per group / backend): a pinned host pair, its
    device copy and the gathered [world, 2] result, so a call is one H2D, one all-gather and
    one D2H instead of a dozen small allocations and synchronising reads"""

    def __init__(self, group):
        import torch
        import torch.distributed as dist
        self.world = dist.get_world_size(group)
        self.nccl = dist.get_backend(group) == "nccl"
        dev = torch.device("cuda", torch.cuda.current_device()) if self.nccl else torch.device("cpu")
        self.host = torch.zeros(2, dtype=torch.float64, pin_memory=self.nccl)
        self.mine = torch.zeros(2, dtype=torch.float64, device=dev)
        self.all = torch.zeros(self.world * 2, dtype=torch.float64, device=dev)   # flat: gloo too
        self.all_host = torch.zeros(self.world * 2, dtype=torch.float64, pin_memory=self.nccl)


_workspaces = {}


def reduce_best(cost, idx, group=None):
    """lexicographic (cost, idx) minimum over ranks == np.argmin's first-minimum rule on the
    concatenated cost vector.  Works on any torch.distributed backend (gloo on CPU, nccl)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(cost), int(idx)
    ws = _workspaces.get(group)
    if ws is None:
        ws = _workspaces[group] = _Workspace(group)
    ws.host[0] = float(cost)
    ws.host[1] = float(idx)
    ws.mine.copy_(ws.host, non_blocking=True)
    dist.all_gather_into_tensor(ws.all, ws.mine, group=group)
    ws.all_host.copy_(ws.all)          # synchronises
    flat = ws.all_host.tolist()
    return min((flat[2 * r], int(flat[2 * r + 1])) for r in range(ws.world))


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


def bind_host_numa(device):
    """One process per GPU: pin this process to the CPUs of the GPU's NUMA node BEFORE allocating
    pinned host buffers (first touch then places them on that node).  Returns the node, or None
    when the platform exposes no placement (f1l_bind_host_numa)."""
    from . import _lib
    node = _lib.lib().f1l_bind_host_numa(int(device))
    return node if node >= 0 else None
