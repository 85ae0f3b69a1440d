"""N>1 host logic on CPU: world_size-2 gloo run of the frame sharding used by bench.py."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jxl_oxide_b200 import sharding  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, num_frames, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.frames_for_rank(num_frames, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    # pretend rank r took (r+1)*10 ms: the step time is the max over ranks
    t = sharding.max_over_ranks_ms(10.0 * (rank + 1))
    dist.barrier()
    if rank == 0:
        out.put((gathered, t))
    dist.destroy_process_group()


@pytest.mark.parametrize("num_frames", [64, 7, 1])
def test_frames_shard_disjoint_and_complete_gloo(num_frames):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, num_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, t = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    flat = sorted(k for part in gathered for k in part)
    assert flat == list(range(num_frames))
    assert all(abs(len(a) - len(b)) <= 1 for a in gathered for b in gathered)
    assert t == 20.0


def test_rank_out_of_range():
    with pytest.raises(ValueError):
        sharding.frames_for_rank(4, 2, 2)


def _gather_worker(rank, world, port, num_frames, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.frames_for_rank(num_frames, rank, world)
    # frame k = a small "packed" u8 image filled with k
    local = [torch.full((4, 6, 3), k, dtype=torch.uint8) for k in mine]
    if num_frames % 2:  # the streaming form: frames handed over as callables that produce them when their round comes
        local = [(lambda t=t: t) for t in local]
    got = sharding.gather_frames(local, num_frames, dst=0)
    dist.barrier()
    if rank == 0:
        out.put([int(t[0, 0, 0]) if t is not None else None for t in got] + [tuple(got[0].shape)])
    else:
        assert got is None
    dist.destroy_process_group()


@pytest.mark.parametrize("num_frames", [8, 7])
def test_gather_frames_in_frame_order_gloo(num_frames):
    """BASELINE config #5's gather (NCCL on the GPUs): rank 0 ends up with every frame, in order, also when the
    last round is ragged."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, world, port, num_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[:-1] == list(range(num_frames))
    assert res[-1] == (4, 6, 3)


def test_gather_frames_single_process_is_identity():
    frames = [torch.zeros(2, 2, 3, dtype=torch.uint8) for _ in range(3)]
    assert sharding.gather_frames(frames, 3) == frames
    got = sharding.gather_frames([(lambda t=t: t) for t in frames], 3)
    assert all(a is b for a, b in zip(got, frames))
