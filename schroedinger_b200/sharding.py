"""Picture- / GOP-parallel sharding across the GPUs of one box.

The picture core has no exchange step: intra pictures, closed GOPs and (frame, reference)
pairs are independent, so ranks never exchange picture data.  torch.distributed is used
only for the benchmark contract's barrier and max-over-ranks timing.
"""
import torch
import torch.distributed as dist


def shard(n_units, rank, world):
    """Indices of the independent units (pictures / GOPs / pairs) rank `rank` owns:
    unit i goes to rank i mod world, the reference's own round-robin over workers."""
    return list(range(rank, n_units, world))


def gop_shard(n_pictures, gop_len, rank, world):
    """Pictures of whole GOPs: a GOP chain (I..P + its references) never crosses GPUs."""
    out = []
    for g in shard((n_pictures + gop_len - 1) // gop_len, rank, world):
        out.extend(range(g * gop_len, min((g + 1) * gop_len, n_pictures)))
    return out


def max_over_ranks(value, device="cpu"):
    """Elapsed time of a multi-rank region = the slowest rank's."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_throughput(units_per_rank, elapsed_max, world):
    """Whole-job throughput under weak scaling: every rank processed units_per_rank."""
    return units_per_rank * world / elapsed_max
