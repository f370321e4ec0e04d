"""
Sharding of the sliding-DFT path across GPUs (one process per GPU, torch.distributed for the plumbing).

Two shardings make sense for this path:

* by CHANNEL -- plans are independent (the whole state lives in the plan, c/src/sdft/sdft.h:175-180),
  so channels are dealt out in contiguous blocks and nothing is ever exchanged;
* by TIME with a 2m-sample halo -- shard g analyses samples [begin_g, end_g) after priming its plan
  with the 2m samples before begin_g (``SDFT.advance``).  Boundaries are multiples of 2m so that every
  shard starts at cursor 0, exactly where the reference's periodic phase restart falls
  (c/src/sdft/sdft.h:566-576).  The only collective is the optional all-gather of synthesized samples.

Nothing here computes: the planners are pure integer arithmetic and ``gather_samples`` is a thin
wrapper over ``torch.distributed.all_gather`` (NCCL on GPUs, gloo in the CPU tests).
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
