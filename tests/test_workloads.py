"""The five BASELINE.json configurations as bench.py / the parity tests see them (raytracer-rust_b200/workloads.py): native
settings per SURVEY.md 8d, object counts per SURVEY.md 8a row a4.  CPU only: scene descriptions, no device."""
import pytest


@pytest.mark.parametrize("name,settings,n_objects", [("C1", (256, 256, 16, 8), 8), ("C2", (800, 600, 256, 30), 5),
                                                     ("C3", (1280, 720, 512, 64), 2), ("C3s", (1280, 720, 512, 64), None),
                                                     ("C4", (1280, 720, 1024, 16), 9)])
def test_named_configs_have_their_native_settings(pt, name, settings, n_objects):
    from raytracer_rust_b200 import workloads
    label, s = workloads.workload(name)
    assert label.startswith(name + " ") and s.settings == settings
    if n_objects is not None:
        assert len(s.objects) == n_objects
    if name == "C3s":
        assert s.sky is not None and sum(len(s.mesh(o.mesh)) for o in s.objects if o.type == pt.OBJ_MESH) == 124840


def test_synthetic_config(pt):
    from raytracer_rust_b200 import workloads
    label, s = workloads.workload("C5", cells=32)
    assert s.settings == (3840, 2160, 256, 16) and "2,048-triangle" in label
    assert len(s.mesh(s.objects[0].mesh)) == 2 * 32 * 32
    with pytest.raises(ValueError):
        workloads.workload("C9")
