"""
Sharding of the sliding-DFT path across GPUs (one process per GPU, torch.distributed for the plumbing).

Two shardings make sense for this path:

* by CHANNEL -- plans are independent (the whole state lives in the plan, c/src/sdft/sdft.h:175-180),
  so channels are dealt out in contiguous blocks and nothing is ever exchanged;
* by TIME with a 2m-sample halo -- shard g analyses samples [begin_g, end_g) after priming its plan
  with the 2m samples before begin_g (``SDFT.advance``).  Boundaries are multiples of 2m so that every
  shard starts at cursor 0, exactly where the reference's periodic phase restart falls
  (c/src/sdft/sdft.h:566-576).  The only collective is the optional all-gather of synthesized samples.

  The halo re-seed reproduces the continuous run up to the reference's own delta-rounding random walk
  (float time domain: ~5e-8 after 2^21 samples) -- fine for float frequency-domain data.
  The EXACT variant for double frequency-domain data exchanges one row of m complex numbers per rank:
  every shard first sums its own accumulator increments (``shard_increment``: a state-only pass), the
  increments are all-gathered (``gather_increments``) and each shard starts from the in-order sum of its
  predecessors' increments (``start_exact``).

Nothing here computes: the planners are pure integer arithmetic, the state helpers only call the plan's
``advance`` / ``state`` / ``set_state``, and the gathers are thin wrappers over
``torch.distributed.all_gather`` (NCCL on GPUs, gloo in the CPU tests).
"""
from dataclasses import dataclass


@dataclass(frozen=True)
class TimeShard:
    rank: int
    begin: int        # first sample analysed by this rank
    end: int          # one past the last sample analysed by this rank
    halo_begin: int   # first sample of the priming window (== begin for rank 0: zero history)

    @property
    def size(self):
        return self.end - self.begin

    @property
    def halo(self):
        return self.begin - self.halo_begin


def time_shards(nsamples, world, dftsize):
    """Splits [0, nsamples) into `world` contiguous shards whose boundaries are multiples of 2*dftsize.
    Trailing ranks may be empty when the signal is shorter than world periods."""
    period = 2 * dftsize
    periods = (nsamples + period - 1) // period
    base, extra = divmod(periods, world)
    shards, begin = [], 0
    for r in range(world):
        n_periods = base + (1 if r < extra else 0)
        end = min(nsamples, begin + n_periods * period)
        shards.append(TimeShard(r, begin, end, max(0, begin - period)))
        begin = end
    return shards


def channel_shards(channels, world):
    """Contiguous [begin, end) channel ranges, sizes differing by at most one."""
    base, extra = divmod(channels, world)
    out, begin = [], 0
    for r in range(world):
        end = begin + base + (1 if r < extra else 0)
        out.append((begin, end))
        begin = end
    return out


def gather_samples(local, shards, group=None):
    """All-gathers the synthesized samples of every time shard into one tensor on every rank.
    `local` is this rank's 1-D tensor of length shards[rank].size."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    longest = max(s.size for s in shards)
    padded = torch.zeros(longest, dtype=local.dtype, device=local.device)
    padded[:local.numel()] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:s.size] for p, s in zip(parts, shards)])


def _halo_history(plan, halo):
    """The 2m samples in front of a shard as plan history (zero history before the signal's start)."""
    import numpy as np
    period = 2 * plan.size
    h = np.zeros(period, dtype=np.float32 if plan.td == "f32" else np.float64)
    halo = np.asarray(halo)
    if halo.size:
        h[period - halo.size:] = halo[-period:]
    return h


def shard_increment(plan, halo, shard):
    """Pass 1 of exact time sharding: what this shard adds to every bin's accumulator (m complex values).
    The plan is left in an undefined state; call ``start_exact`` before analysing the shard."""
    import numpy as np
    plan.reset()
    plan.set_state(0, history=_halo_history(plan, halo),
                   accumulators=np.zeros(plan.size, np.complex64 if plan.fd == "f32" else np.complex128))
    plan.advance(shard)
    return plan.state()[2]


def gather_increments(increment, group=None):
    """All-gathers the per-shard accumulator increments; returns a (world, m) complex array on every rank."""
    import numpy as np
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    local = torch.view_as_real(torch.from_numpy(np.ascontiguousarray(increment)))
    if dist.get_backend(group) == "nccl":
        local = local.cuda()
    parts = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(parts, local.contiguous(), group=group)
    return np.stack([torch.view_as_complex(p.cpu().contiguous()).numpy() for p in parts])


def start_exact(plan, halo, increments, rank):
    """Pass 2: puts the plan exactly where a continuous run would be at the start of shard `rank`:
    history = the 2m samples before it, accumulators = the preceding shards' increments added in order."""
    import numpy as np
    acc = np.zeros(plan.size, np.complex64 if plan.fd == "f32" else np.complex128)
    for g in range(rank):
        acc = acc + increments[g].astype(acc.dtype)
    plan.reset()
    plan.set_state(0, history=_halo_history(plan, halo), accumulators=acc)
