"""Multi-GPU: one process per GPU, rays sharded by image, one all-gather.

The reference scales with nn.DataParallel: scatter on the batch dimension,
replicate the module, gather the outputs on GPU 0 (run.py:636-644).  The render
path has no cross-image arithmetic, so the B200 version
is: every rank renders its own contiguous slice of the batch with the fused
kernels, then ONE NCCL all-gather of the packed [rgb(3), depth, mask] tiles
brings the full batch to every rank (``gather=True``); inversion needs no
gradient collective because latents and poses are per image.
"""

import torch
import torch.distributed as dist


def shard_range(batch, world_size, rank):
    """Contiguous [start, stop) of images for ``rank``; remainders go to the
    lowest ranks (same split nn.DataParallel's scatter produces)."""
    if not (0 <= rank < world_size):
        raise ValueError('rank %d outside world of %d' % (rank, world_size))
    base, rem = divmod(batch, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_inputs(tensors, world_size, rank):
    """Slices every [B, ...] tensor (None passes through) to this rank's images."""
    batch = next(t.shape[0] for t in tensors.values() if t is not None)
    a, b = shard_range(batch, world_size, rank)
    return {k: (None if t is None else t[a:b]) for k, t in tensors.items()}


def pack_outputs(rgb, depth, mask):
    """[b,H,W,3], [b,H,W], [b,H,W] -> [b,H,W,5] (one buffer, one collective)."""
    return torch.cat((rgb, depth.unsqueeze(-1), mask.unsqueeze(-1)), dim=-1)


def unpack_outputs(packed):
    return packed[..., :3], packed[..., 3], packed[..., 4]


def all_gather_outputs(rgb, depth, mask, batch, group=None):
    """All ranks end up with the full-batch (rgb, depth, mask).

    Shards may be ragged (batch not divisible by the world size): every rank
    pads its packed tile to the largest shard, gathers, and trims.
    """
    world = dist.get_world_size(group)
    packed = pack_outputs(rgb, depth, mask)
    sizes = [shard_range(batch, world, r) for r in range(world)]
    longest = max(b - a for a, b in sizes)
    if packed.shape[0] < longest:
        pad = packed.new_zeros((longest - packed.shape[0],) + tuple(packed.shape[1:]))
        packed = torch.cat((packed, pad), dim=0)
    out = [torch.empty_like(packed) for _ in range(world)]
    dist.all_gather(out, packed.contiguous(), group=group)
    full = torch.cat([o[: b - a] for o, (a, b) in zip(out, sizes)], dim=0)
    return unpack_outputs(full)


def render_sharded(render_fn, batch_inputs, batch, gather=True, group=None):
    """Runs ``render_fn(**shard)`` on this rank's images; ``render_fn`` returns
    (rgb, depth, mask, ...).  With ``gather`` the full batch comes back."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    shard = shard_inputs(batch_inputs, world, rank)
    rgb, depth, mask = render_fn(**shard)[:3]
    if world == 1 or not gather:
        return rgb, depth, mask
    return all_gather_outputs(rgb, depth, mask, batch, group)
