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


def gather_frames(local_frames, num_frames, dst=0, group=None):
    """BASELINE config #5's delivery step: every rank holds the packed frames `frames_for_rank()` gave it (equally
    shaped tensors, on the GPU with the nccl backend); rank `dst` receives all `num_frames` of them in frame order, the
    other ranks get None. One `gather` per round of `world_size` frames (frame k of round r lives on rank k): with
    nccl the copies go GPU to GPU over NVLink / NVSwitch, nothing is staged through the host. Frames are
    independent, so this is the path's only exchange and it carries finished pixels, never intermediate data."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    mine = frames_for_rank(num_frames, rank, world)
    if len(local_frames) != len(mine):
        raise ValueError(f"rank {rank} holds {len(local_frames)} frames, its share of {num_frames} is {len(mine)}")
    if world == 1:
        return [f() if callable(f) else f for f in local_frames]
    rounds = (num_frames + world - 1) // world
    # a rank without a frame in the last (ragged) round still takes part in the collective with a dummy
    proto = local_frames[0] if local_frames else None
    if callable(proto):
        proto = None
    out = [None] * num_frames
    for r in range(rounds):
        have = r < len(local_frames)
        if have:
            f = local_frames[r]
            if callable(f):  # a frame still being decoded: f() blocks until it is packed (decode overlaps earlier rounds)
                f = f()
            send = f.contiguous()
        else:
            if proto is None and have is False and local_frames and callable(local_frames[0]):
                proto = local_frames[0]()
            if proto is None:
                raise ValueError("a rank without any frame cannot size its placeholder; pass at least world_size frames")
            send = torch.empty_like(proto)
        recv = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
        dist.gather(send, recv, dst=dst, group=group)
        if rank == dst:
            for src in range(world):
                k = r * world + src
                if k < num_frames:
                    out[k] = recv[src]
    return out if rank == dst else None
