"""Data-parallel sharding of independent images (reference generate_conditional.py:317-340, 557-569).

Every image carries its own covariance state, mask, measurement and noise, so there is no exchange step in the
sampling loop: rank r of W samples the global image indices r, r+W, r+2W, ... and the only collective is the final
gather of the finished images (NCCL on the GPUs, gloo in the CPU tests).  The reference seeds masks and measurement
noise from (seed, rank) (generate_conditional.py:317), which makes results depend on the world size; here every
per-image random input is derived from (seed, global image index) so that any disjoint cover of 0..N-1 gives the
same images.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_indices(n_images, rank, world_size):
    """Global image indices of this rank: rank, rank + W, ...  (a disjoint cover of range(n_images))."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    return list(range(rank, n_images, world_size))


def image_seed(seed, index):
    """Seed of the per-image generators (mask, measurement noise, initial noise): a function of the global index only."""
    return int((np.uint64(seed) * np.uint64(1_000_003) + np.uint64(index)) % np.uint64(2 ** 31 - 1))


def gather_images(local_images, local_indices, n_images):
    """All-gather of the finished images: `local_images` (n_local, C, H, W) with their global `local_indices`.
    Returns (n_images, C, H, W) on every rank, ordered by global index.  Ranks may hold different counts."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out = torch.empty((n_images,) + tuple(local_images.shape[1:]), dtype=local_images.dtype, device=local_images.device)
        out[torch.as_tensor(local_indices, dtype=torch.long, device=local_images.device)] = local_images
        return out
    world = dist.get_world_size()
    n_max = (n_images + world - 1) // world
    shape = (n_max,) + tuple(local_images.shape[1:])
    pad = torch.zeros(shape, dtype=local_images.dtype, device=local_images.device)
    pad[: local_images.shape[0]] = local_images
    idx = torch.full((n_max,), -1, dtype=torch.long, device=local_images.device)
    idx[: len(local_indices)] = torch.as_tensor(local_indices, dtype=torch.long, device=local_images.device)
    bufs = [torch.empty_like(pad) for _ in range(world)]
    ibufs = [torch.empty_like(idx) for _ in range(world)]
    dist.all_gather(bufs, pad)
    dist.all_gather(ibufs, idx)
    out = torch.empty((n_images,) + tuple(local_images.shape[1:]), dtype=local_images.dtype, device=local_images.device)
    seen = torch.zeros(n_images, dtype=torch.bool, device=local_images.device)
    for b, i in zip(bufs, ibufs):
        keep = i >= 0
        out[i[keep]] = b[keep]
        seen[i[keep]] = True
    if not bool(seen.all()):
        raise RuntimeError("gather_images: the ranks' index sets do not cover all images")
    return out
