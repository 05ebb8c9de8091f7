"""Host logic of the multi-GPU path on CPU: index sharding, world-size independent seeds, and the final gather over a
world_size-2 gloo group (the product uses the same code over NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from free_hunch_b200.outputs import aggregate_mean, decode_uint8, image_filename, psnr_uint8
from free_hunch_b200.sharding import gather_images, image_seed, shard_indices


def test_shard_indices_cover():
    for n in (0, 1, 7, 8, 3000):
        for w in (1, 2, 4, 8):
            parts = [shard_indices(n, r, w) for r in range(w)]
            flat = sorted(i for p in parts for i in p)
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert shard_indices(3000, 3, 8)[:3] == [3, 11, 19] and len(shard_indices(3000, 3, 8)) == 375
    with pytest.raises(ValueError):
        shard_indices(4, 2, 2)


def test_image_seed_depends_on_index_only():
    seeds = {image_seed(0, i) for i in range(3000)}
    assert len(seeds) == 3000
    assert image_seed(0, 17) == image_seed(0, 17) != image_seed(1, 17)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = shard_indices(n, rank, world)
    # the "sample" of image i is a deterministic function of its global seed only
    imgs = torch.stack([torch.from_numpy(np.random.RandomState(image_seed(0, i)).standard_normal((3, 4, 4))) for i in idx]) \
        if idx else torch.zeros((0, 3, 4, 4), dtype=torch.float64)
    out = gather_images(imgs, idx, n)
    agg = aggregate_mean(10.0 + rank)                 # per-rank means 10 and 11 -> 10.5 on every rank
    q.put((rank, out.numpy(), agg))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [5, 8])
def test_gather_world2_gloo(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    res = {r: a for r, a, _ in got}
    assert all(abs(g - 10.5) < 1e-6 for _, _, g in got)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = np.stack([np.random.RandomState(image_seed(0, i)).standard_normal((3, 4, 4)) for i in range(n)])
    for r in range(2):
        assert np.array_equal(res[r], ref)          # same images in global order on every rank, as with world size 1


def test_gather_single_process():
    imgs = torch.arange(3 * 2, dtype=torch.float32).reshape(3, 2)
    out = gather_images(imgs, [2, 0, 1], 3)
    assert torch.equal(out, imgs[[1, 2, 0]])


def test_output_side_helpers():
    """decode (encoders.py:72-73: scale, clip, truncate), file names, PSNR (skimage definition)."""
    x = torch.tensor([-1.5, -1.0, -0.999, 0.0, 0.5, 0.999, 1.0, 2.0])
    assert decode_uint8(x).tolist() == [0, 0, 0, 128, 191, 255, 255, 255] and decode_uint8(x.double()).dtype == torch.uint8
    assert image_filename(17, 3) == "000017_000003.png"
    a = torch.zeros(3, 8, 8, dtype=torch.uint8)
    b = a.clone(); b[0, 0, 0] = 16
    assert psnr_uint8(a, a) == float("inf")
    assert abs(psnr_uint8(a, b) - 10 * np.log10(255.0 ** 2 / (256.0 / 192))) < 1e-9
    assert aggregate_mean(3.25) == 3.25                # no process group: the value itself
