"""Bounded, seeded runs of the randomized parity cases (tests/fuzz_lib.py) on the GPU: single images under
AUTO and with the streaming kernels forced, device-resident batches with odd pitches, YUV420 frames."""
import numpy as np
import pytest

import fuzz_lib
import libiqo_b200 as iqo

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path,seed", [(iqo.PATH_AUTO, 101), (iqo.PATH_STREAM, 102), (iqo.PATH_NO_STREAM, 103), (iqo.PATH_MMA, 104)])
def test_fuzz_single_images(path, seed):
    rng = np.random.RandomState(seed)
    stats, bad, done = {}, [], 0
    for _ in range(200):
        case = fuzz_lib.single_case(rng)
        res = fuzz_lib.run_single(rng, case, path)
        if res is None:
            continue
        done += 1
        stats[res[0]] = stats.get(res[0], 0) + 1
        if not res[1]:
            bad.append((case, res[0]))
    assert not bad, bad[:5]
    assert done >= 120
    if path == iqo.PATH_STREAM:   # every streaming family was exercised
        for k in ("lanczos_stream", "ratio_stream"):
            assert stats.get(k, 0) > 0, stats
        assert any(k.endswith("_stream") and k.startswith("half") for k in stats), stats
    if path == iqo.PATH_AUTO:     # the Area / Linear streaming kernels take single images too
        assert stats.get("area_down", 0) > 0 and any(k.startswith("linear_") and k != "linear_mma" for k in stats), stats
    if path == iqo.PATH_MMA:      # both tensor-path variants
        assert stats.get("lanczos_mma", 0) > 0 and stats.get("lanczos_mma_dp2a", 0) > 0, stats


@pytest.mark.parametrize("path,seed", [(iqo.PATH_AUTO, 201), (iqo.PATH_STREAM, 202), (iqo.PATH_MMA, 203)])
def test_fuzz_device_batches(path, seed):
    pytest.importorskip("torch")
    rng = np.random.RandomState(seed)
    bad, done = [], 0
    for _ in range(80):
        case = fuzz_lib.batch_case(rng)
        res = fuzz_lib.run_batch(rng, case, path)
        if res is None:
            continue
        done += 1
        if not res[1]:
            bad.append((case, res[0]))
    assert not bad, bad[:5]
    assert done >= 60


@pytest.mark.parametrize("path,seed", [(iqo.PATH_AUTO, 301), (iqo.PATH_STREAM, 302), (iqo.PATH_MMA, 303)])
def test_fuzz_row_bands(path, seed):
    pytest.importorskip("torch")
    rng = np.random.RandomState(seed)
    bad, done = [], 0
    for _ in range(120):
        res = fuzz_lib.run_band_case(rng, path)
        if res is None:
            continue
        done += 1
        if not res[1]:
            bad.append(res)
    assert not bad, bad[:5]
    assert done >= 90


def test_fuzz_yuv420_frames():
    pytest.importorskip("torch")
    rng = np.random.RandomState(301)
    bad, done = [], 0
    for _ in range(40):
        case = fuzz_lib.yuv_case(rng)
        res = fuzz_lib.run_yuv(rng, case)
        if res is None:
            continue
        done += 1
        if not res[1]:
            bad.append(case)
    assert not bad, bad[:5]
    assert done >= 25
