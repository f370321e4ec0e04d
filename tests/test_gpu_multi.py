"""
Multi-GPU paths on real devices (skipped on a single-GPU box): one process per GPU, NCCL for the plumbing.

* config 3 -- a chirp cut into time shards with a 2m-sample halo: every rank primes a fresh plan with its
  halo (`SDFT.advance`), runs the fused analysis+synthesis on its shard, and the synthesized samples are
  all-gathered (`sdft_b200.shard.gather_samples`).  Rank 0 compares with the continuous CPU reference run.
* config 4 -- independent channels dealt out in contiguous blocks, no data-path collective at all; rank 0
  gathers the synthesized samples only to compare them with the oracle.
"""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_path):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    import torch.distributed as dist
    from sdft_b200 import SDFT, workloads
    from sdft_b200.shard import channel_shards, gather_samples, time_shards
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))

    # ---- time shards (config 3 shape, reduced length) ----
    n, m = 1 << 19, 2048
    x = workloads.chirp(n)
    shards = time_shards(n, world, m)
    s = shards[rank]
    plan = SDFT(m, "hann", 0.5, td="f32", fd="f32")
    xd = torch.from_numpy(x[s.halo_begin:s.end]).cuda()
    if s.halo:
        plan.advance(xd[:s.halo])
    y_local = plan.roundtrip(xd[s.halo:])
    y_time = gather_samples(y_local, shards)

    # ---- channel shards (config 4 shape, reduced) ----
    channels, nc, mc = 6, 20000, 1024
    a, b = channel_shards(channels, world)[rank]
    xc = np.stack([workloads.channel_noise(c, nc) for c in range(a, b)])
    batch = SDFT(mc, "hann", 1, td="f32", fd="f64", channels=b - a)
    yc = batch.roundtrip(torch.from_numpy(xc).cuda())
    parts = [torch.empty((hi - lo, nc), dtype=torch.float32, device="cuda") for lo, hi in channel_shards(channels, world)]
    dist.all_gather(parts, yc.reshape(b - a, nc))          # equal shard sizes here (6 channels, world 2)
    # ---- exact time shards of a double frequency-domain plan: one all-gather of accumulator increments ----
    from sdft_b200.shard import gather_increments, shard_increment, start_exact
    ne, me = 1 << 17, 512
    xe = workloads.white_noise(ne, seed=78)
    se = time_shards(ne, world, me)[rank]
    pe = SDFT(me, "hann", 1, td="f32", fd="f64")
    inc = gather_increments(shard_increment(pe, xe[se.halo_begin:se.begin], xe[se.begin:se.end]))
    start_exact(pe, xe[se.halo_begin:se.begin], inc, rank)
    rows_e = torch.from_numpy(pe.sdft(xe[se.begin:se.begin + 32])).cuda()
    rows_all = [torch.empty_like(rows_e) for _ in range(world)]
    dist.all_gather(rows_all, rows_e)
    if rank == 0:
        np.savez(out_path, y_time=y_time.cpu().numpy(), y_chan=torch.cat(parts).cpu().numpy(),
                 rows_exact=torch.stack(rows_all).cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_time_and_channel_shards(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from oracle import Oracle
    from sdft_b200 import workloads
    out = str(tmp_path / "multi.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)

    n, m = 1 << 19, 2048
    x = workloads.chirp(n)
    want = Oracle("f32", "f32", m, "hann", 0.5).roundtrip(x)
    assert got["y_time"].shape == want.shape
    assert np.abs(got["y_time"].astype(np.float64) - want).max() <= 1e-3 * np.abs(want).max()
    delay = int((m - 1) * 0.5)
    assert abs(workloads.snr_db(x, got["y_time"], delay) - workloads.snr_db(x, want, delay)) < 0.01

    ne, me = 1 << 17, 512
    xe = workloads.white_noise(ne, seed=78)
    walk, pos = Oracle("f32", "f64", me, "hann", 1.0), 0
    from sdft_b200.shard import time_shards
    for s in time_shards(ne, 2, me):
        walk.advance(xe[pos:s.begin])
        pos = s.begin
        want_rows = walk.clone().sdft(xe[s.begin:s.begin + 32])
        assert np.abs(got["rows_exact"][s.rank] - want_rows).max() <= 1e-9 * np.abs(want_rows).max(), s.rank

    for c in range(6):
        xc = workloads.channel_noise(c, 20000)
        yo = Oracle("f32", "f64", 1024, "hann", 1.0).roundtrip(xc)
        assert np.abs(got["y_chan"][c] - yo).max() <= 2e-6, c
