"""Frame → rank assignment for multi-GPU decode (one process per GPU).

The decode path shards by frame (independent frames, no data-path collective), the
way the reference's CLI renders keyframes in a `par_iter` (crates/jxl-oxide-cli/src/decode.rs:293-301).
`torch.distributed` is used only for the barrier and for reducing the step time (max over ranks).
"""


def frames_for_rank(num_frames, rank, world_size):
    """Round-robin: frame k goes to rank k % world_size (BASELINE config #5: 64 frames over 8 ranks)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return list(range(rank, num_frames, world_size))


def max_over_ranks_ms(local_ms, device=None):
    """The time of a step is the slowest rank's device time."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(local_ms)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
