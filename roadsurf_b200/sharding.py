"""Point sharding for multi-GPU runs: one process per GPU, contiguous equal blocks of the point
index, no collective on the data path (points are independent for the whole run).
torch.distributed is used only for the barrier and for the max-over-ranks timing reduction."""


def shard_range(npoints, rank, world_size):
    """Points [a, b) owned by `rank`: GPU g gets [g*P/G, (g+1)*P/G)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank outside world")
    a = npoints * rank // world_size
    b = npoints * (rank + 1) // world_size
    return a, b


def reduce_over_ranks(value, op="max", device="cpu"):
    """max / sum of a python float over all ranks (identity when torch.distributed is not
    initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return float(t.item())


def barrier(device=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if device is not None and str(device).startswith("cuda"):
            dist.barrier(device_ids=[int(str(device).split(":")[1])])
        else:
            dist.barrier()
