"""The N > 1 host-side logic on CPU: two processes over gloo exercise the
sharding rules and the id/timing plumbing that bench.py uses under torchrun
(the GPU all-reduce itself is covered by the 2-GPU bench run)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from recur_b200 import dist as rdist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r, w, lr = rdist.env_rank()
    assert (r, w, lr) == (rank, world, rank)
    payload = bytes(range(128)) if rank == 0 else b"\0" * 128
    got = rdist.broadcast_bytes(dist, payload, 128, 0)
    slowest = rdist.max_over_ranks(dist, 1.5 + rank)
    lo, hi = rdist.shard_bounds(2_000_001, rank, world)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, got == bytes(range(128)), slowest, lo, hi))


def test_two_ranks_over_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _, _ in res)
    assert all(s == 2.5 for _, _, s, _, _ in res)
    (_, _, _, lo0, hi0), (_, _, _, lo1, hi1) = res
    assert lo0 == 0 and hi0 == lo1 and hi1 - lo1 == hi0 - lo0 == 1_000_000


def test_shards_are_disjoint_and_equal():
    from recur_b200 import dist as rdist
    for world in (1, 2, 4, 8):
        bounds = [rdist.shard_bounds(2_000_000, r, world) for r in range(world)]
        assert all(b[1] - b[0] == 2_000_000 // world for b in bounds)
        assert all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))


def test_stream_positions_follow_rnn_char_epoch_spacing():
    from recur_b200 import dist as rdist
    length, n = 1001, 8
    spacing = (length - 1) // n
    for step in (0, 5, 999):
        pos = rdist.stream_positions(length, n, step)
        for j, p in enumerate(pos):
            off = step + j * spacing
            if off >= length - 1:
                off -= length - 1
            assert p == off % (length - 1)


def test_reference_arm_prints_the_contract_line():
    """bench.py --impl reference on a tiny sample (CPU only)."""
    import json
    import subprocess
    import oracle
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    env = dict(os.environ, RECUR_BENCH_TINY="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                         env=env, timeout=600)
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "chars/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0
