"""
Host-side logic of the multi-GPU path, on CPU: the shard planners, and a world_size-2 gloo run in which
each rank analyses its time shard (primed with the 2m-sample halo) and the synthesized samples are
all-gathered.  The per-rank transform is done by the ORACLE here (this is a test of the sharding logic
and the collective, not of the kernels; the same flow runs on GPUs in bench/tests with sdft_b200.SDFT).
"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sdft_b200.shard import channel_shards, time_shards


def test_time_shards_cover_signal_on_period_boundaries():
    for n, world, m in [(2 ** 16, 8, 2048), (100000, 4, 1000), (5000, 8, 1024), (4096, 2, 2048), (1, 3, 8)]:
        shards = time_shards(n, world, m)
        assert len(shards) == world
        assert shards[0].begin == 0 and shards[-1].end == n
        for a, b in zip(shards, shards[1:]):
            assert a.end == b.begin
        for s in shards:
            assert s.begin % (2 * m) == 0 or s.begin == n
            assert s.halo == (0 if s.begin == 0 else min(2 * m, s.begin))
            assert 0 <= s.size


def test_channel_shards_are_balanced():
    for ch, world in [(512, 8), (5, 2), (3, 8), (64, 1)]:
        parts = channel_shards(ch, world)
        assert parts[0][0] == 0 and parts[-1][1] == ch
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1
        for (a0, b0), (a1, b1) in zip(parts, parts[1:]):
            assert b0 == a1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, m, out_path):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import Oracle
    from sdft_b200.shard import gather_samples, time_shards
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x = np.random.default_rng(123).uniform(-1, 1, n).astype(np.float32)
    shards = time_shards(n, world, m)
    s = shards[rank]
    plan = Oracle("f32", "f32", m, "hann", 0.5)
    if s.halo:
        plan.sdft(x[s.halo_begin:s.begin])          # prime: rows discarded (SDFT.advance on the GPU)
    y = plan.isdft(plan.sdft(x[s.begin:s.end])) if s.size else np.zeros(0, np.float32)
    full = gather_samples(torch.from_numpy(y), shards)
    if rank == 0:
        np.save(out_path, full.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_time_sharded_roundtrip_with_gloo(tmp_path):
    from oracle import Oracle
    n, m, world = 20000, 256, 2
    out = str(tmp_path / "y.npy")
    mp.spawn(_worker, args=(world, _free_port(), n, m, out), nprocs=world, join=True)
    got = np.load(out)
    x = np.random.default_rng(123).uniform(-1, 1, n).astype(np.float32)
    ref = Oracle("f32", "f32", m, "hann", 0.5)
    want = ref.isdft(ref.sdft(x))
    assert got.shape == want.shape
    # rank 0 is bit-identical; rank 1 is re-seeded from its halo: float-FD re-seed noise ~4e-6
    half = time_shards(n, world, m)[0].end
    assert np.array_equal(got[:half], want[:half])
    assert np.abs(got[half:] - want[half:]).max() <= 1e-4 * np.abs(want).max()


def _inc_worker(rank, world, port, out_path):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from sdft_b200.shard import gather_increments
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(rank)
    mine = (rng.uniform(-1, 1, 19) + 1j * rng.uniform(-1, 1, 19)).astype(np.complex128)
    got = gather_increments(mine)
    if rank == 1:
        np.save(out_path, got)
    dist.barrier()
    dist.destroy_process_group()


def test_gather_increments_with_gloo(tmp_path):
    """The one exchange step of exact time sharding: every rank ends up with all ranks' accumulator increments."""
    out = str(tmp_path / "inc.npy")
    mp.spawn(_inc_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    for r in range(2):
        rng = np.random.default_rng(r)
        want = (rng.uniform(-1, 1, 19) + 1j * rng.uniform(-1, 1, 19)).astype(np.complex128)
        assert np.array_equal(got[r], want)


def test_c_abi_planners_agree_with_python():
    """sdft_b200_time_shard / sdft_b200_channel_shard (for C and C++ callers) are the same arithmetic as the
    Python planners; pure host functions, no device needed."""
    import ctypes
    import random
    from sdft_b200 import _lib
    from sdft_b200.shard import channel_shards, time_shards
    lib = _lib.load()
    rnd = random.Random(7)
    sz = ctypes.c_size_t
    for _ in range(300):
        n, world, m = rnd.randrange(0, 1 << 20), rnd.randrange(1, 17), rnd.choice([1, 3, 37, 512, 1000, 2048])
        for s in time_shards(n, world, m):
            b, e, h = sz(), sz(), sz()
            assert lib.sdft_b200_time_shard(n, world, m, s.rank, ctypes.byref(b), ctypes.byref(e), ctypes.byref(h)) == 0
            assert (b.value, e.value, h.value) == (s.begin, s.end, s.halo_begin)
        ch = rnd.randrange(0, 600)
        for r, (lo, hi) in enumerate(channel_shards(ch, world)):
            b, e = sz(), sz()
            assert lib.sdft_b200_channel_shard(ch, world, r, ctypes.byref(b), ctypes.byref(e)) == 0
            assert (b.value, e.value) == (lo, hi)
    assert lib.sdft_b200_time_shard(100, 0, 8, 0, None, None, None) != 0
    assert lib.sdft_b200_channel_shard(8, 2, 2, None, None) != 0
