"""ptc_multi_*: the in-process multi-GPU path of the C ABI (one host thread per GPU, one NCCL reduce of the film).
Needs two GPUs; on a one-GPU box only the single-device degenerate case runs."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENES = os.path.join(ROOT, "scenes")
pytestmark = pytest.mark.gpu


def _scene(pt):
    s = pt.load_scene_from_json(os.path.join(SCENES, "semesterbild.json"))
    return s, s.to_core().commit(0)


def test_multi_single_device_equals_render(pt):
    s, cs = _scene(pt)
    st = s.render_settings(width=120, height=90, spp=6, max_depth=8, seed=2)
    whole, ws = cs.render(s.camera, st)
    for shard in (pt.SHARD_SAMPLES, pt.SHARD_TILES):
        img, ms = cs.multi([0]).render(s.camera, st, shard)
        assert np.array_equal(img, whole) and ms.rays == ws.rays and ms.paths == ws.paths
    packed, _ = cs.multi([0]).render_u32(s.camera, st)
    want, _ = cs.render_u32(s.camera, st)
    assert np.array_equal(packed, want)
    with pytest.raises(pt.PtcError):
        cs.multi([0, 0])  # duplicate device
    with pytest.raises(pt.PtcError):
        cs.multi([1] if pt.device_count() > 1 else [7])  # devices[0] must be the scene's device / out of range


@pytest.mark.skipif("__import__('ptload').load().device_count() < 2")
def test_multi_two_devices_shards_are_invisible(pt):
    s, cs = _scene(pt)
    n = min(pt.device_count(), 4)
    m = cs.multi(list(range(n)))
    st = s.render_settings(width=200, height=150, spp=8, max_depth=30, seed=5)
    whole, ws = cs.render(s.camera, st)
    for shard in (pt.SHARD_SAMPLES, pt.SHARD_TILES):
        img, ms = m.render(s.camera, st, shard)
        assert ms.paths == ws.paths == 200 * 150 * 8 and ms.rays == ws.rays
        # the films are reduced as 64-bit integers: the multi-GPU image IS the one-GPU image
        assert np.array_equal(img, whole)
    # more devices than samples: sample sharding leaves devices idle, the image is still the whole image
    st1 = s.render_settings(width=64, height=48, spp=1, max_depth=4, seed=5)
    a, _ = cs.render(s.camera, st1)
    b, _ = m.render(s.camera, st1, pt.SHARD_SAMPLES)
    assert np.array_equal(a, b)
