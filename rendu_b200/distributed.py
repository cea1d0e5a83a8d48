"""Multi-GPU plumbing of PathTracer::render over torch.distributed (one process per GPU; NCCL on the B200 box, gloo
in the CPU tests). The scene is replicated on every rank; a frame is split either by image tile (disjoint row bands,
no reduction: the bands are gathered on rank 0) or by stratification cell (every rank renders every pixel for
sid % world == rank; the per-pixel sums are combined with ONE reduce, then normalised on rank 0). SURVEY.md §8e.

The functions only move tensors: what renders the local share is the caller's business (the CUDA path in bench.py, a
stand-in in the gloo tests).
"""
import torch
import torch.distributed as dist

DEFAULT_TILE_ROWS = 2  # band height of the tile split: 2 rows balance the C3 frame to 1 % over 8 ranks (8 rows: 6 %; tools/tile_balance.py)


def tile_spec(rank, world, tile_rows=DEFAULT_TILE_ROWS):
    """(tile_rows, tile_index, tile_count) for rb_render_opts: bands of `tile_rows` rows dealt round-robin."""
    return (int(tile_rows), int(rank), int(world))


def sample_spec(rank, world):
    """(sample_index, sample_count) for rb_render_opts: stratification cells with sid % world == rank."""
    return (int(rank), int(world))


def local_rows(height, rank, world, tile_rows=DEFAULT_TILE_ROWS):
    """Image rows rendered by `rank` under the tile split, in the order the renderer enumerates them."""
    ys = torch.arange(height)
    return ys[((ys // tile_rows) % world) == rank]


def combine_tile_split(local_band, height, rank, world, tile_rows=DEFAULT_TILE_ROWS, dst=0, group=None):
    """local_band: (rows_of_this_rank, W, C) final pixels of this rank's rows. Returns the full (H, W, C) image on `dst`
    (None elsewhere). Disjoint tiles: a gather, not a reduction; bands are padded to a common row count."""
    if world == 1:
        return local_band
    counts = [int(local_rows(height, r, world, tile_rows).numel()) for r in range(world)]
    pad = max(counts)
    w, c = local_band.shape[1], local_band.shape[2]
    send = torch.zeros((pad, w, c), dtype=local_band.dtype, device=local_band.device)
    send[: local_band.shape[0]] = local_band
    bufs = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    full = torch.empty((height, w, c), dtype=local_band.dtype, device=local_band.device)
    for r in range(world):
        rows = local_rows(height, r, world, tile_rows).to(local_band.device)
        full[rows] = bufs[r][: counts[r]]
    return full


def combine_sample_split(local_sum, rank, world, dst=0, group=None):
    """local_sum: (H, W, 3) per-pixel SUMS of clamped sample radiance over this rank's cells. One reduce(SUM) to `dst`
    (fp32, W*H*3 values: 24.9 MB at 1080p); valid because the reference clamps per sample BEFORE accumulating
    (PathTracer.cpp:288), so the estimator is a plain sum. Returns the total on `dst`, None elsewhere."""
    if world > 1:
        dist.reduce(local_sum, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return local_sum if rank == dst else None
