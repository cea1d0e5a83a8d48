"""World-size-2 gloo tests (CPU) of the multi-GPU plumbing: tile split gathers disjoint bands into the exact
single-process image; sample split reduces per-rank sums into the single-process sums (fp32 order aside).
The local renderer is the ORACLE here (there is no GPU in this container) — the CUDA path is exercised under NCCL by
bench.py --gpus N (torchrun) on the B200 box and by tests/test_pathtracer_gpu.py for the split options themselves."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rendu_b200 import distributed as rd

W, H, S, D = 40, 28, 16, 4


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _scene():
    from oracle import oracle as orc
    from rendu_b200 import workloads as wl
    return orc.OracleScene(wl.tiny_scene()), wl.tiny_camera(W / H)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc, cam = _scene()
    full, _ = sc.render(cam, W, H, S, D, seed=4, rng="counter", threads=1)
    # tile split: this rank's rows only (the oracle renders the whole frame; the band is sliced out, which is what a
    # renderer restricted to those rows produces because every pixel depends only on (seed, pixel, sid))
    rows = rd.local_rows(H, rank, world, tile_rows=4)
    band = torch.from_numpy(full)[rows].contiguous()
    tiled = rd.combine_tile_split(band, H, rank, world, tile_rows=4)
    # sample split: sums over this rank's stratification cells
    part, _ = sc.render(cam, W, H, S, D, seed=4, rng="counter", threads=1, sample_split=rd.sample_spec(rank, world))
    total = rd.combine_sample_split(torch.from_numpy(part * np.float32(S)), rank, world)
    if rank == 0:
        np.save(out + "_tiled.npy", tiled.numpy())
        np.save(out + "_sum.npy", total.numpy())
        np.save(out + "_full.npy", full)
    dist.barrier()
    dist.destroy_process_group()


def test_tile_and_sample_split_world2(tmp_path):
    out = str(tmp_path / "r")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    full = np.load(out + "_full.npy")
    assert np.array_equal(np.load(out + "_tiled.npy").view(np.uint32), full.view(np.uint32))
    np.testing.assert_allclose(np.load(out + "_sum.npy") / S, full, rtol=0, atol=2e-6)


def test_row_assignment_covers_every_row_once():
    for h, world, tr in ((1080, 8, 8), (1080, 4, 16), (37, 3, 4), (5, 8, 8)):
        seen = torch.zeros(h, dtype=torch.int32)
        for r in range(world):
            rows = rd.local_rows(h, r, world, tr)
            seen[rows] += 1
            assert rd.tile_spec(r, world, tr) == (tr, r, world)
        assert (seen == 1).all()
