"""Host-side logic of the multi-GPU path on CPU: world_size-2 `gloo` process group (rendezvous on 127.0.0.1)."""
import os
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeLib:
    """stands in for libveritas_b200 where libnccl cannot hand out an id (CPU box): a recognisable 128-byte pattern"""

    def vrt_nccl_unique_id(self, buf):
        for i in range(128):
            buf[i] = (7 * i + 3) % 256
        return 0


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    from veritas_b200.parallel import slab_bounds, broadcast_unique_id, max_over_ranks
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nx = 4096
        x0, x1 = slab_bounds(nx, rank, world)
        # the slabs tile the domain in rank order
        t = torch.tensor([x0, x1], dtype=torch.int64)
        parts = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(parts, t)
        bounds = [tuple(p.tolist()) for p in parts]
        uid = broadcast_unique_id(dist, FakeLib(), rank)
        slow = max_over_ranks(dist, 10.0 + rank)
        q.put((rank, bounds, uid, slow))
    finally:
        dist.destroy_process_group()


def test_slab_partition_and_id_broadcast_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = bytes((7 * i + 3) % 256 for i in range(128))
    for rank, bounds, uid, slow in out:
        assert bounds == [(0, 2048), (2048, 4096)]
        assert uid == expect            # every rank holds rank 0's id
        assert slow == 11.0             # max over ranks


def test_slab_bounds_errors():
    sys.path.insert(0, ROOT)
    from veritas_b200.parallel import slab_bounds
    assert slab_bounds(262144, 3, 8) == (98304, 131072)
    with pytest.raises(ValueError):
        slab_bounds(100, 0, 8)
    with pytest.raises(ValueError):
        slab_bounds(32, 0, 8)
    with pytest.raises(ValueError):
        slab_bounds(64, 8, 8)


def test_reference_arm_runs_on_rank0_only():
    """bench.py --impl reference under torchrun: ranks other than 0 exit 0 without work or output."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
