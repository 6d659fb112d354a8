"""Host-side sharding helpers for multi-GPU runs (one process per GPU).

Nothing in the resize path reduces across devices, so sharding is pure index arithmetic:
frames split into contiguous blocks, a huge image splits into destination row bands whose
source rows (band + halo) come from the planner.  No CUDA call is made here; the functions are
exercised on CPU with a world_size-2 gloo group in tests/test_sharding.py.
"""
from . import plan_query


def frame_shard(n_frames, world, rank):
    """Contiguous block of frames of `rank`: (first, count).  Blocks differ by at most one frame."""
    base, extra = divmod(n_frames, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def band_shard(dst_h, world, rank):
    """Destination row band of `rank`: (y0, rows)."""
    return frame_shard(dst_h, world, rank)


def band_source_rows(kind, degree, src_w, src_h, dst_w, dst_h, px_scale, y0, rows):
    """Source rows [s0, s0+n) that destination rows [y0, y0+rows) need (band + halo), from the
    planner's first-tap map; identical to iqo_cuda_band_src_rows but needs no device."""
    if rows <= 0:
        return 0, 0
    q = plan_query(kind, degree, src_w, src_h, dst_w, dst_h, px_scale, 1)
    first, n = q["first"], q["numCoefs"]
    lo = min(max(int(first[y0]), 0), src_h - 1)
    hi = min(max(int(first[y0 + rows - 1]) + n - 1, 0), src_h - 1)
    return lo, hi - lo + 1


def max_over_ranks(value, dist=None, device=None):
    """MAX all-reduce of a python float (timing aggregation); identity without a process group."""
    if dist is None or not dist.is_initialized():
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
