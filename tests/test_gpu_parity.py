"""Closest-hit / BSDF / RNG parity of the CUDA core (libptcore.so, through the C ABI) against the oracle, on a B200.
Same cases as test_hostsim_parity.py, at larger sizes."""
import pytest

import parity_cases as pc

pytestmark = pytest.mark.gpu
KIND = "gpu"


@pytest.mark.parametrize("name,over", [("semesterbild.json", {}),                                   # C2 at native 800x600
                                       ("cornell-box/scene.json", dict(width=256, height=256)),      # C1
                                       ("veach-mis/scene.json", {})])                                # C4 at native 1280x720
def test_primary_rays_shipped_scenes(pt, name, over):
    pc.check_primary_rays_shipped_scene(KIND, pt, name, over)


def test_random_rays_composite_scene(pt):
    pc.check_random_rays_composite(KIND, pt, n=400000)


def test_many_objects(pt):
    pc.check_many_objects(KIND, pt, n=200000)


def test_mesh_build_facts(pt):
    pc.check_mesh_build_facts(KIND, pt)


def test_axis_aligned_and_degenerate_rays(pt):
    pc.check_axis_aligned_and_degenerate_rays(KIND, pt)


def test_scatter_all_materials(pt):
    # CUDA libm vs glibc: sampled directions agree to 1e-5, weights to 1e-4 relative (stated tolerance)
    pc.check_scatter_all_materials(KIND, pt, n=20000, tol_dir=1e-5, tol_att=1e-4)


def test_philox_and_resolve(pt):
    pc.check_philox_and_resolve(KIND, pt)


def test_misses_and_ids(pt):
    pc.check_sky_lookup(KIND, pt)


def test_all_dead_mesh(pt):
    pc.check_all_dead_mesh(KIND, pt)


def test_small_meshes(pt):
    pc.check_small_meshes(KIND, pt)


def test_synthetic_heightfield(pt):
    # C5's generator at 256x256 cells = 131,072 triangles (the full 2 M-triangle build is exercised by the bench tools)
    pc.check_synthetic_heightfield(KIND, pt, cells=256, n=200000)


def test_synthetic_heightfield_full_size(pt):
    # BASELINE config C5's mesh at its full size: 1000x1000 cells = 2,000,000 triangles (299 k wide nodes, 96 MB of
    # triangles); primary rays + 300 k random rays, every hit record bit-identical to the oracle's
    r, stats = pc.check_synthetic_heightfield(KIND, pt, cells=1000, n=300000)
    assert stats.mesh_rays > 0 and stats.nodes_visited / stats.mesh_rays > 4
