"""Multi-GPU host logic: contiguous world shards and the one collective of the path — an allgather
of per-rollout costs followed by an identical arg-min / top-k on every rank.

Worlds are independent (the reference already steps two unrelated ensembles per SimulationStep,
/root/reference/eggshell/model.cc:67-68), so there is no data-path exchange inside a horizon."""
import numpy as np


def shard_range(n_worlds_total, rank, world_size):
    """Contiguous slice [lo, hi) of the global world index space owned by `rank`."""
    base, rem = divmod(int(n_worlds_total), int(world_size))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def allgather_costs(local_costs):
    """local_costs: 1-D torch tensor (CPU/gloo or CUDA/nccl).  Returns the concatenation over ranks
    in rank order.  Uneven shards are padded to the longest one."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_costs.clone()
    ws = dist.get_world_size()
    n = torch.tensor([local_costs.numel()], dtype=torch.int64, device=local_costs.device)
    sizes = [torch.zeros_like(n) for _ in range(ws)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes)
    if all(s == m for s in sizes):
        out = torch.empty(ws * m, dtype=local_costs.dtype, device=local_costs.device)
        dist.all_gather_into_tensor(out, local_costs.contiguous())
        return out
    pad = torch.full((m,), float("inf"), dtype=local_costs.dtype, device=local_costs.device)
    pad[: local_costs.numel()] = local_costs
    out = torch.empty(ws * m, dtype=local_costs.dtype, device=local_costs.device)
    dist.all_gather_into_tensor(out, pad)
    return torch.cat([out[r * m: r * m + sizes[r]] for r in range(ws)])


def select_best(all_costs, k=1):
    """Deterministic top-k (lowest cost, ties by lowest global world index) — identical on every rank."""
    c = np.asarray(all_costs.detach().cpu().numpy() if hasattr(all_costs, "detach") else all_costs, dtype=np.float64)
    order = np.lexsort((np.arange(c.size), c))
    return order[:k], c[order[:k]]


def owner_of(world_index, n_worlds_total, world_size):
    """(rank, local index) of a global world index under shard_range."""
    for r in range(world_size):
        lo, hi = shard_range(n_worlds_total, r, world_size)
        if lo <= world_index < hi:
            return r, world_index - lo
    raise IndexError(world_index)
