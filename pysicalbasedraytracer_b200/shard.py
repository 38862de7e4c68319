"""Multi-GPU sharding of one render: the scene is replicated on every GPU, the per-pixel
sample-index range [s0, s1) is split into contiguous sub-ranges, one per rank (every (pixel, sample) is
addressable in O(1): HaltonSampler::GetIndexForSample, Sampler/Halton.cpp:61-81), each rank accumulates
a private fp32 RGBA *sum* film (A = sample count) and the films are summed by ONE reduce to rank 0.
There is no other data-path exchange: path tracing has no cross-sample dependency.

On the GPUs all of this happens INSIDE the C-ABI (pbrgpu_init_multi / pbrgpu_comm_attach: sharding + ncclReduce + resolve,
csrc/pbrgpu.cu); this module is the host-side twin used by launchers that keep the films in torch tensors (film.py) and by
the world_size-2 gloo tests that cover the N > 1 logic on CPU.  Plumbing only; no rendering happens here."""
import torch
import torch.distributed as dist


def shard_samples(s0, s1, rank, world):
    """Contiguous, balanced split of [s0, s1): the first (n % world) ranks get one extra sample.  This IS the split the
    library applies inside pbrgpu_render on a multi-GPU context (pbrgpu_shard_samples, csrc/pbrgpu.cu): one definition."""
    from . import api
    return api.shard_samples(s0, s1, rank, world)


def reduce_film(accum, dst=0, group=None):
    """Sum the per-rank RGBA sum films into rank `dst` (in place).  Fixed reduction topology => run-to-run
    reproducible.  Returns the tensor (meaningful on `dst` only)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return accum


def resolve_mean(accum):
    """RGBA sums -> per-pixel mean RGB (divide by the sample count carried in A)."""
    return accum[..., :3] / accum[..., 3:4].clamp_min(1.0)
