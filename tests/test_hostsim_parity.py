"""The product's device headers (csrc/pt_*.h), compiled by g++ into tests/hostsim, against the oracle — the closest
this GPU-less container gets to running the kernels.  The same cases run on the B200 in test_gpu_parity.py."""
import pytest

import parity_cases as pc

KIND = "sim"


@pytest.mark.parametrize("name,over", [("semesterbild.json", dict(width=400, height=300)),
                                       ("cornell-box/scene.json", dict(width=128, height=128)),
                                       ("veach-mis/scene.json", dict(width=320, height=180))])
def test_primary_rays_shipped_scenes(pt, name, over):
    pc.check_primary_rays_shipped_scene(KIND, pt, name, over)


def test_random_rays_composite_scene(pt):
    pc.check_random_rays_composite(KIND, pt, n=120000)


def test_many_objects(pt):
    pc.check_many_objects(KIND, pt, n=20000)


def test_mesh_build_facts(pt):
    pc.check_mesh_build_facts(KIND, pt)


def test_axis_aligned_and_degenerate_rays(pt):
    pc.check_axis_aligned_and_degenerate_rays(KIND, pt)


def test_scatter_all_materials(pt):
    # same libm on both sides here, so the sampled directions agree to the last bit
    pc.check_scatter_all_materials(KIND, pt, n=5000, tol_dir=0.0, tol_att=0.0)


def test_philox_and_resolve(pt):
    pc.check_philox_and_resolve(KIND, pt)


def test_misses_and_ids(pt):
    pc.check_sky_lookup(KIND, pt)


def test_all_dead_mesh(pt):
    pc.check_all_dead_mesh(KIND, pt)


def test_small_meshes(pt):
    pc.check_small_meshes(KIND, pt)


def test_synthetic_heightfield(pt):
    pc.check_synthetic_heightfield(KIND, pt, cells=48, n=40000)
