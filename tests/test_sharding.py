"""CPU tests of the multi-GPU host logic with a world_size-2 gloo group: frame blocks and
row bands (+halo) computed per rank, each rank's shard resized by the oracle, results gathered
and compared with the unsharded oracle.  (The GPU path of the same sharding is
tests/test_gpu_parity.py::test_row_bands_equal_whole_image / test_multi_device_drivers...)"""
import os
import socket

import numpy as np
import pytest

from libiqo_b200 import sharding
from oracle_lib import AREA, LANCZOS, LINEAR, lcg_image, oracle_resize

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def test_frame_and_band_shards_partition_exactly():
    for n in (1, 7, 8, 4096, 12000):
        for world in (1, 2, 3, 8):
            parts = [sharding.frame_shard(n, world, r) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            for (a, c), (b, _) in zip(parts, parts[1:]):
                assert a + c == b
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def test_band_source_rows_cfg5():
    # cfg5: 32768 -> 12000, Lanczos4 (22 taps): 1500-row bands need about 4096 + 22 source rows
    for rank in range(8):
        y0, rows = sharding.band_shard(12000, 8, rank)
        s0, sn = sharding.band_source_rows(LANCZOS, 4, 32768, 32768, 12000, 12000, 1, y0, rows)
        assert rows == 1500 and 4096 <= sn <= 4096 + 24
        assert s0 >= 0 and s0 + sn <= 32768


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, case, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    kind, deg, px, sw, sh, dw, dh, n_frames = case
    # ---- row bands of one image: every rank only touches its band + halo rows ----
    src = lcg_image(sh, sw, seed=3)
    y0, rows = sharding.band_shard(dh, world, rank)
    s0, sn = sharding.band_source_rows(kind, deg, sw, sh, dw, dh, px, y0, rows)
    rc, whole = oracle_resize(kind, src, dw, dh, deg, px)
    assert rc == 0
    # a band computed from the whole image equals the band of the whole result by construction;
    # what the sharding must get right is that [s0, s0+sn) covers every row the band reads:
    poisoned = src.copy()
    poisoned[:s0] = 0xEE
    poisoned[s0 + sn:] = 0x11
    rc, band = oracle_resize(kind, poisoned, dw, dh, deg, px)
    objs = [None] * world
    dist.all_gather_object(objs, band[y0:y0 + rows].copy())   # bands may be ragged
    parts = [torch.from_numpy(o) for o in objs]
    stitched = torch.cat(parts).numpy()
    assert np.array_equal(stitched, whole), "row-band sharding lost halo rows"
    # ---- frame blocks of a batch ----
    f0, cnt = sharding.frame_shard(n_frames, world, rank)
    local = [oracle_resize(kind, lcg_image(sh, sw, seed=100 + f), dw, dh, deg, px)[1] for f in range(f0, f0 + cnt)]
    objs = [None] * world
    dist.all_gather_object(objs, (f0, [a.tobytes() for a in local]))
    frames = {}
    for first, blobs in objs:
        for i, b in enumerate(blobs):
            frames[first + i] = b
    assert sorted(frames) == list(range(n_frames))
    for f in (0, n_frames - 1):
        assert frames[f] == oracle_resize(kind, lcg_image(sh, sw, seed=100 + f), dw, dh, deg, px)[1].tobytes()
    # ---- timing aggregation: max over ranks ----
    assert sharding.max_over_ranks(1.0 + rank, dist) == float(world)
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(out_dir, "ok%d" % rank), "w").close()


@pytest.mark.parametrize("case", [(LANCZOS, 4, 1, 256, 192, 94, 71, 5), (AREA, 0, 1, 120, 90, 40, 31, 3),
                                  (LINEAR, 0, 1, 40, 30, 100, 75, 4)])
def test_world_size_2_gloo(tmp_path, case):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), case, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(world))
