"""
bench_sharded.py -- the two shardings BASELINE.json names, measured by `bench.py --gpus N` on every N:

  c3_time_sharded     configs[2]: the 2^26-sample chirp, m = 2048, float FD, latency 0.5, cut into N time shards
                      with a 2m-sample halo (sdft_b200.shard.time_shards).  Rank r primes a fresh plan with the
                      4096 samples in front of its shard (`advance`), runs the fused analysis+synthesis on its
                      shard, and the synthesized samples are all-gathered over NCCL.  STRONG scaling: the job
                      is the same 2^26 samples whatever N is; rank 0 also runs the whole signal on its own GPU
                      in the same process, which gives `speedup_vs_n1` and the N = 1 result the sharded one is
                      compared with (samples, reconstruction SNR per python/examples/latency.py:30-56, and the
                      analysis rows right after every shard boundary against a continuous run).
  c3_exact_f64        the same split for a DOUBLE frequency-domain plan on a 2^22 prefix: the halo re-seed
                      misses the reference's float-delta random walk, so every shard first sums
                      its own accumulator increments (a state-only pass), one all-gather of m complex values
                      per rank distributes them, and each shard starts from the in-order sum of its
                      predecessors' (sdft_b200.shard.shard_increment / gather_increments / start_exact).
  c4_channel_sharded  configs[3]: 512 independent channels x 2^20 samples, m = 1024, double FD, 512/N channels
                      per rank in ONE batched plan, no data-path collective.  STRONG scaling again.

Reference anchors: the plan holds all state (c/src/sdft/sdft.h:175-180), which is what makes channels
independent; the modulation phase restarts every 2m samples (sdft.h:566-576), which is why shard boundaries sit on
multiples of 2m (cursor 0).  Nothing here touches oracle/: the in-line checks compare GPU runs with GPU runs
(sharded against continuous); the oracle-pinned versions of the same comparisons live in tests/.
"""
import ctypes
import math

import numpy as np

C3_N, C3_M, C3_LATENCY = 1 << 26, 2048, 0.5
C3E_N = 1 << 22
C4_CHANNELS, C4_N, C4_M = 512, 1 << 20, 1024
PROBE_ROWS = 64


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _chirp_device(torch, dev, n_total, begin, count):
    """sdft_b200.workloads.chirp evaluated on the device (phase in float64, cast to float32)."""
    t = torch.arange(begin, begin + count, dtype=torch.float64, device=dev)
    return torch.sin((math.pi * 0.25 / float(n_total)) * t * t).to(torch.float32)


def _snr_db(torch, x, y, delay):
    xd = x[: x.numel() - delay].double()
    e = y[delay:].double() - xd
    return float(10 * torch.log10((xd * xd).mean() / (e * e).mean()))


def _checked(res, checks):
    """In-line parity: every check is (what, passed).  A leg whose check fails is reported as an ERROR (with
    its numbers kept for diagnosis) -- a wrong result must not read as a measurement.  No exception is raised
    here: the other ranks are already waiting in the leg's closing barrier."""
    failed = [what for what, ok in checks if not ok]
    res["parity_ok"] = not failed
    if failed:
        res["error"] = "parity check failed: " + "; ".join(failed)
    return res


class _Comm:
    """barrier / max-over-ranks / timing helpers shared by the legs"""

    def __init__(self, torch, dist, dev, rank, world):
        self.torch, self.dist, self.dev, self.rank, self.world = torch, dist, dev, rank, world

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, x, op="max"):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return float(t.item())

    def timed(self, fn, reps, setup=None):
        """min over `reps` of (max over ranks of the CUDA-event time of fn), barrier on both sides"""
        torch = self.torch
        best = None
        for _ in range(reps):
            if setup:
                setup()
            self.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            t = self.reduce(e0.elapsed_time(e1) * 1e-3)
            self.barrier()
            best = t if best is None else min(best, t)
        return best


def _gather_equal(torch, dist, local, world):
    out = torch.empty(world * local.numel(), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local)
    return out


def c3_time_sharded(torch, dist, SDFT, comm, reps=3, n=C3_N):
    from sdft_b200.shard import gather_samples, time_shards
    rank, world, dev = comm.rank, comm.world, comm.dev
    m, delay = C3_M, int((C3_M - 1) * C3_LATENCY)
    shards = time_shards(n, world, m)
    s = shards[rank]
    equal = len({sh.size for sh in shards}) == 1
    xs = _chirp_device(torch, dev, n, s.halo_begin, s.end - s.halo_begin)
    halo, mine = xs[: s.halo], xs[s.halo:]
    y_local = torch.empty_like(mine)
    plan = SDFT(m, "hann", C3_LATENCY, td="f32", fd="f32")
    plan._use_torch_stream()
    f_adv, f_rt = plan._f("advance"), plan._f("roundtrip_n")
    launches0 = plan.launches

    def shard_job():
        if s.halo:
            f_adv(plan._h, s.halo, _ptr(halo))
        if s.size:
            f_rt(plan._h, s.size, _ptr(mine), _ptr(y_local))

    shard_job()
    plan._check()
    t_compute = comm.timed(shard_job, reps, setup=plan.reset)
    plan._check()
    launches_per_job = (plan.launches - launches0) // (reps + 1)

    # the one collective of the path: all-gather of the synthesized samples
    gathered = None
    t_gather, gather_bytes = 0.0, 0
    if world > 1:
        def gather():
            nonlocal gathered
            gathered = _gather_equal(torch, dist, y_local, world) if equal else gather_samples(y_local, shards)
        gather()
        t_gather = comm.timed(gather, reps)
        gather_bytes = n * 4
    else:
        gathered = y_local

    # rows right after this rank's shard boundary: halo-primed plan against a continuous run from t = 0
    rows_diff = 0.0
    if s.halo and s.size:
        k = min(PROBE_ROWS, s.size)
        prefix = _chirp_device(torch, dev, n, 0, s.begin + k)
        cont = SDFT(m, "hann", C3_LATENCY, td="f32", fd="f32")
        cont.advance(prefix[: s.begin])
        rows_c = cont.sdft(prefix[s.begin:])
        probe = SDFT(m, "hann", C3_LATENCY, td="f32", fd="f32")
        probe.advance(halo)
        rows_s = probe.sdft(mine[:k])
        rows_diff = float((rows_s - rows_c).abs().max() / rows_c.abs().max())
        del prefix, rows_c, rows_s, cont, probe
    rows_diff = comm.reduce(rows_diff)

    # N = 1 in the same process: rank 0 runs the whole signal on its GPU (the others idle at the barrier)
    res = {}
    if rank == 0:
        x_all = xs if world == 1 else _chirp_device(torch, dev, n, 0, n)
        y_one = torch.empty_like(x_all)
        one = SDFT(m, "hann", C3_LATENCY, td="f32", fd="f32")
        one._use_torch_stream()

        def whole():
            one._f("roundtrip_n")(one._h, n, _ptr(x_all), _ptr(y_one))
        whole()
        one._check()
        t_one = None
        for _ in range(reps):
            one.reset()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            whole()
            e1.record()
            torch.cuda.synchronize()
            t = e0.elapsed_time(e1) * 1e-3
            t_one = t if t_one is None else min(t_one, t)
        one._check()
        snr_n, snr_1 = _snr_db(torch, x_all, gathered, delay), _snr_db(torch, x_all, y_one, delay)
        diff = float((gathered.double() - y_one.double()).abs().max())
        scale = float(y_one.abs().max())
        t_job = t_compute + t_gather
        res = {
            "workload": "configs[2]: 2^26-sample chirp 0 -> 0.25 fs, m=2048, f32 TD / f32 FD, hann, latency 0.5; %d time "
                        "shard(s) on multiples of 2m with a 2m-sample halo, fused analysis+synthesis per shard, "
                        "all-gather of the synthesized samples; chirp generated on the device (formula of "
                        "sdft_b200.workloads.chirp)" % world,
            "scaling": "strong", "n_samples": n, "shard_samples": [sh.size for sh in shards],
            "samples_per_s": n / t_job, "bin_updates_per_s": n * m / t_job,
            "compute_ms": t_compute * 1e3, "allgather_ms": t_gather * 1e3,
            "allgather": {"collective": "ncclAllGather of the synthesized samples (torch.distributed "
                                        "all_gather_into_tensor)" if world > 1 else None,
                          "bytes_total": gather_bytes, "bytes_per_rank": gather_bytes // world,
                          "GBps": (gather_bytes / t_gather / 1e9) if t_gather > 0 else None},
            "n1_ms": t_one * 1e3, "speedup_vs_n1": t_one / t_job,
            "snr_db": snr_n, "snr_db_n1": snr_1,
            "max_abs_diff_vs_n1": diff, "max_rel_diff_vs_n1": diff / scale,
            "rows_after_boundary_max_rel_diff_vs_continuous": rows_diff,
            "gpu_launches_per_shard": int(launches_per_job),
            "checks": "in-line (a failure turns this leg into an error): |snr - snr_n1| <= 0.01 dB, rows <= 1e-4, samples <= 1e-3 of full scale",
        }
        res = _checked(res, [("|snr - snr_n1| = %.4f dB > 0.01" % abs(snr_n - snr_1), abs(snr_n - snr_1) <= 0.01),
                             ("rows after a shard boundary %.3g > 1e-4" % rows_diff, rows_diff <= 1e-4),
                             ("samples %.3g > 1e-3 of full scale" % (diff / scale), diff <= 1e-3 * scale)])
    comm.barrier()
    return res


def c3_exact_f64(torch, dist, SDFT, comm, reps=3, n=C3E_N):
    """Exact time sharding of a float-TD / double-FD plan: increments all-gathered, rows to 1e-9."""
    from sdft_b200.shard import gather_increments, gather_samples, shard_increment, start_exact, time_shards
    rank, world, dev = comm.rank, comm.world, comm.dev
    m = C3_M
    shards = time_shards(n, world, m)
    s = shards[rank]
    x_all = _chirp_device(torch, dev, C3_N, 0, n)        # the first 2^22 samples of the config-3 chirp
    halo_np = x_all[s.halo_begin:s.begin].cpu().numpy()
    mine = x_all[s.begin:s.end]
    plan = SDFT(m, "hann", C3_LATENCY, td="f32", fd="f64")
    plan._use_torch_stream()

    # pass 1: this shard's accumulator increments (state-only kernel), then ONE all-gather of m complex values
    inc = shard_increment(plan, halo_np, mine)
    torch.cuda.synchronize()
    t_pass1 = comm.timed(lambda: plan._f("advance")(plan._h, s.size, _ptr(mine)), reps)
    t_inc, incs = 0.0, inc[None]
    if world > 1:
        incs = gather_increments(inc)
        local = torch.view_as_real(torch.from_numpy(np.ascontiguousarray(inc))).to(dev).contiguous()

        def gather_inc():
            _gather_equal(torch, dist, local.view(-1), world)
        gather_inc()
        t_inc = comm.timed(gather_inc, reps)
    # pass 2: start exactly where the continuous run would be, fused round trip over the shard
    y_local = torch.empty_like(mine)

    def pass2():
        plan._f("roundtrip_n")(plan._h, s.size, _ptr(mine), _ptr(y_local))
    t_pass2 = comm.timed(pass2, reps, setup=lambda: start_exact(plan, halo_np, incs, rank))
    plan._check()
    gathered = gather_samples(y_local, shards) if world > 1 else y_local

    # rows right after the boundary: exact start against the continuous run, and the plain halo re-seed for scale
    k = min(PROBE_ROWS, s.size)
    cont = SDFT(m, "hann", C3_LATENCY, td="f32", fd="f64")
    if s.begin:
        cont.advance(x_all[: s.begin])
    rows_c = cont.sdft(mine[:k])
    start_exact(plan, halo_np, incs, rank)
    rows_e = plan.sdft(mine[:k])
    exact_diff = float((rows_e - rows_c).abs().max() / rows_c.abs().max())
    halo_diff = 0.0
    if s.halo:
        h = SDFT(m, "hann", C3_LATENCY, td="f32", fd="f64")
        h.advance(x_all[s.halo_begin:s.begin])
        halo_diff = float((h.sdft(mine[:k]) - rows_c).abs().max() / rows_c.abs().max())
    exact_diff, halo_diff = comm.reduce(exact_diff), comm.reduce(halo_diff)

    res = {}
    if rank == 0:
        one = SDFT(m, "hann", C3_LATENCY, td="f32", fd="f64")
        y_one = one.roundtrip(x_all)
        torch.cuda.synchronize()
        diff = float((gathered.double() - y_one.double()).abs().max())
        res = {
            "workload": "first 2^22 samples of the config-3 chirp, m=2048, f32 TD / f64 FD, hann, latency 0.5; %d exact "
                        "time shard(s): state-only pass -> all-gather of the accumulator increments -> fused round "
                        "trip from the exact state" % world,
            "n_samples": n,
            "pass1_increments_ms": t_pass1 * 1e3, "pass2_roundtrip_ms": t_pass2 * 1e3,
            "increments_allgather_ms": t_inc * 1e3,
            "increments_allgather": {"collective": "ncclAllGather of m complex128 accumulator increments per rank"
                                     if world > 1 else None, "bytes_per_rank": m * 16, "bytes_total": m * 16 * world},
            "rows_after_boundary_max_rel_diff_vs_continuous": exact_diff,
            "rows_with_plain_halo_reseed_for_comparison": halo_diff,
            "max_abs_diff_vs_n1": diff,
            "checks": "in-line (a failure turns this leg into an error): rows <= 1e-9 of full scale, samples <= 2e-6",
        }
        res = _checked(res, [("rows after a shard boundary %.3g > 1e-9" % exact_diff, exact_diff <= 1e-9),
                             ("samples %.3g > 2e-6" % diff, diff <= 2e-6)])
    comm.barrier()
    return res


def c4_channel_sharded(torch, dist, SDFT, comm, reps=3, channels=C4_CHANNELS, n=C4_N):
    from sdft_b200 import workloads
    from sdft_b200.shard import channel_shards
    rank, world, dev = comm.rank, comm.world, comm.dev
    m = C4_M
    a, b = channel_shards(channels, world)[rank]
    mine = b - a
    x = torch.empty((mine, n), dtype=torch.float32, device=dev)
    for c in range(a, b):
        x[c - a] = torch.from_numpy(workloads.channel_noise(c, n)).to(dev)
    y = torch.empty_like(x)
    plan = SDFT(m, "hann", 1, td="f32", fd="f64", channels=mine)
    plan._use_torch_stream()
    launches0 = plan.launches

    def job():
        plan._f("roundtrip_n")(plan._h, n, _ptr(x), _ptr(y))
    job()
    plan._check()
    launches_per_job = plan.launches - launches0
    t_job = comm.timed(job, reps, setup=plan.reset)
    plan._check()

    # the rows themselves at the HBM rate, on as many samples per channel as fit a bounded tile
    free_b, _ = torch.cuda.mem_get_info()
    n_rows = 4096
    while mine * n_rows * m * 16 > min(0.5 * free_b, 48 << 30) and n_rows > 256:
        n_rows //= 2
    tile = torch.empty((mine, n_rows, m), dtype=torch.complex128, device=dev)
    rows_plan = SDFT(m, "hann", 1, td="f32", fd="f64", channels=mine)
    rows_plan._use_torch_stream()
    xr = x[:, :n_rows].contiguous()

    def rows_job():
        rows_plan._f("sdft_batch")(rows_plan._h, n_rows, _ptr(xr), _ptr(tile))
    rows_job()
    t_rows = comm.timed(rows_job, reps)
    rows_plan._check()
    rows_bytes = channels * n_rows * m * 16
    del tile

    # every channel is an independent plan: first and last channel of this rank against single-channel plans
    worst = 0.0
    for c in sorted({0, mine - 1}):
        single = SDFT(m, "hann", 1, td="f32", fd="f64")
        ys = single.roundtrip(x[c])
        worst = max(worst, float((ys.double() - y[c].double()).abs().max()))
    worst = comm.reduce(worst)
    checksum = comm.reduce(float(y.double().sum()), op="sum")
    energy = comm.reduce(float((y.double() ** 2).sum()), op="sum")
    res = {}
    if rank == 0:
        res = {
            "workload": "configs[3]: 512 channels x 2^20 samples (default_rng([0x5DF70004, c])), m=1024, f32 TD / f64 FD, "
                        "hann; %d channels per rank in one batched plan, fused analysis+synthesis, no data-path "
                        "collective" % (channels // world),
            "scaling": "strong", "channels": channels, "channels_per_rank": channels // world, "n_samples": n,
            "bin_updates_per_s": channels * n * m / t_job, "samples_per_s": channels * n / t_job, "ms": t_job * 1e3,
            "rows_path": {"what": "sdft_batch: (channels, %d, m) complex128 rows into device memory, one launch" % n_rows,
                          "GBps_all_ranks": rows_bytes / t_rows / 1e9, "GBps_per_gpu": rows_bytes / t_rows / 1e9 / world,
                          "bin_updates_per_s": channels * n_rows * m / t_rows},
            "max_abs_diff_vs_single_channel_plan": worst,
            "checksum_sum_y": checksum, "checksum_sum_y2": energy,
            "gpu_launches_per_job": int(launches_per_job),
            "checks": "in-line (a failure turns this leg into an error): batched channels equal single-channel plans to 2e-6; the checksums are the "
                      "same numbers at every N",
        }
        res = _checked(res, [("batched vs single-channel plan %.3g > 2e-6" % worst, worst <= 2e-6)])
    comm.barrier()
    return res


def run_all(torch, dist, SDFT, dev, rank, world, reps=3):
    """Returns {"c3_time_sharded": ..., "c3_exact_f64": ..., "c4_channel_sharded": ...} on rank 0, {} elsewhere;
    a failing leg reports its error instead of costing the headline line."""
    comm = _Comm(torch, dist, dev, rank, world)
    out = {}
    for name, fn in (("c3_time_sharded", c3_time_sharded), ("c3_exact_f64", c3_exact_f64),
                     ("c4_channel_sharded", c4_channel_sharded)):
        err = None
        try:
            res = fn(torch, dist, SDFT, comm, reps=reps)
        except Exception as exc:
            res, err = {}, "%s: %s" % (type(exc).__name__, exc)
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        # a rank that failed must not leave the others hanging in a collective of the next leg
        failed = comm.reduce(1.0 if err else 0.0)
        if rank == 0:
            out[name] = res if not failed else {"error": err or "a rank other than 0 failed"}
        if failed and world > 1:
            break
    return out
