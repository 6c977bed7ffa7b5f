"""N>1 host logic on CPU: world_size-2 gloo run of the pair-sharding + output gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import rife46 as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_run(frames, tasks, fr):
    lo, hi = fr
    for p, _ in tasks:
        assert lo <= p and p + 1 < hi  # a rank only touches its own frame range (one-frame halo)
    if not tasks:
        return torch.zeros((0,) + tuple(frames.shape[1:3]) + (3,))
    return torch.stack([(1 - t) * frames[p, ..., :3] + t * frames[p + 1, ..., :3] for p, t in tasks])


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import __graft_entry__ as ge
    ge.load_package()
    from cfi_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = O.synthetic_clip(6, 16, 24, seed=3)
    tasks, _ = O.build_tasks(6, [3, 2, 1, 4], ([4], True))  # ragged: 2+1+0+3+skip = 6 tasks over 5 pairs
    out = shard.interpolate_sharded(_fake_run, frames, tasks, dist)
    # chunked compute-then-gather pipeline (the N>1 device path of bench.py): same result as the single gather
    slices = shard.shard_tasks(len(tasks), world)
    counts = [b - a for a, b in slices]
    lo, hi = slices[rank]
    mine = list(tasks[lo:hi])
    local = torch.full((max(counts), 16, 24, 3), float("nan"))

    def run_slice(a, b):
        local[a:b] = _fake_run(frames, mine[a:b], shard.frame_range(mine))

    bufs = shard.forward_and_gather(run_slice, local, counts, dist, nchunks=2)
    if rank == 0:
        # numpy arrays pickle by value; torch tensors would travel as shared-memory handles that die with this process
        q.put((out.numpy(), torch.cat([b_[:c] for b_, c in zip(bufs, counts)], 0).numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_equals_single_process(pkg):
    from cfi_b200 import shard
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, got_pipelined = (torch.from_numpy(a) for a in q.get(timeout=120))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    frames = O.synthetic_clip(6, 16, 24, seed=3)
    tasks, _ = O.build_tasks(6, [3, 2, 1, 4], ([4], True))
    want = _fake_run(frames, tasks, shard.frame_range(tasks))
    assert torch.equal(got, want)
    assert torch.equal(got_pipelined, want)


def test_shard_tasks_balance(pkg):
    from cfi_b200 import shard
    assert shard.shard_tasks(63, 8) == [(0, 8), (8, 16), (16, 24), (24, 32), (32, 40), (40, 48), (48, 56), (56, 63)]
    assert shard.shard_tasks(3, 4) == [(0, 1), (1, 2), (2, 3), (3, 3)]
    assert shard.shard_tasks(0, 2) == [(0, 0), (0, 0)]
    assert shard.frame_range([(2, 0.5), (2, 0.75), (5, 0.5)]) == (2, 7)
    assert shard.chunk_bounds(63, 4) == [(0, 16), (16, 32), (32, 48), (48, 63)]
    assert shard.chunk_bounds(2, 4) == [(0, 1), (1, 2)]
