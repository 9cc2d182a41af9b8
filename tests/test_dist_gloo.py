"""world_size-2 gloo run of the multi-GPU host logic (raytracer-rust_b200/dist.py) on CPU: sharding by sample range
and by interleaved tile, one SUM reduce of the film, 1/spp scaling.  The per-rank renderer here is the oracle (Philox
mode, which honours sample ranges); on the B200 the same code path is fed by ptc_render_accumulate."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H, SPP, DEPTH, SEED = 40, 36, 6, 6, 5


def _worker(rank, world, port, mode, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ptload
    pt = ptload.load()
    from bindings import OracleScene
    import importlib
    pdist = importlib.import_module("raytracer_rust_b200.dist")
    scene = pt.load_scene_from_json(os.path.join(ROOT, "scenes", "cornell-box", "scene.json"))
    orc = OracleScene(scene)

    def render_fn(b, e, tile_mod, tile_rem):
        img, _ = orc.render(scene.camera, W, H, SPP, DEPTH, seed=SEED, sample_begin=b, sample_end=e, threads=2)
        film = img * SPP  # the oracle returns sum / spp; the film the ranks exchange is the sum
        if tile_mod > 0:  # keep only this rank's interleaved 32x32 tiles
            tiles_x = (W + 31) // 32
            ys, xs = np.mgrid[0:H, 0:W]
            tile = (ys // 32) * tiles_x + xs // 32
            film = film * (tile % tile_mod == tile_rem)[..., None]
        return torch.from_numpy(np.ascontiguousarray(film, dtype=np.float32).reshape(-1))

    out = pdist.render_sharded(render_fn, SPP, mode=mode)
    if rank == 0:
        np.save(out_path, out.numpy())
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["samples", "tiles"])
def test_two_rank_render_equals_single(pt, tmp_path, mode):
    from bindings import OracleScene
    port = 29500 + (os.getpid() % 2000) + (0 if mode == "samples" else 1)
    out_path = str(tmp_path / f"film_{mode}.npy")
    mp.spawn(_worker, args=(2, port, mode, out_path), nprocs=2, join=True)
    got = np.load(out_path).reshape(H, W, 3)
    scene = pt.load_scene_from_json(os.path.join(ROOT, "scenes", "cornell-box", "scene.json"))
    want, _ = OracleScene(scene).render(scene.camera, W, H, SPP, DEPTH, seed=SEED)
    assert np.allclose(got, want, rtol=1e-5, atol=1e-6)


def test_sample_ranges_partition():
    import importlib
    import ptload
    ptload.load()
    pdist = importlib.import_module("raytracer_rust_b200.dist")
    for spp in (1, 7, 256, 1000):
        for world in (1, 2, 3, 8):
            r = [pdist.sample_range(spp, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == spp
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1
    assert pdist.shard("tiles", 16, 3, 8) == (0, 16, 8, 3)
