"""Import helper: the package directory keeps the reference's hyphen (`raytracer-rust_b200/`), which Python cannot
import by name.  `ptload.load()` returns it as the module `raytracer_rust_b200`."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))


def load():
    name = "raytracer_rust_b200"
    if name in sys.modules:
        return sys.modules[name]
    path = os.path.join(_ROOT, "raytracer-rust_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location(name, path, submodule_search_locations=[os.path.dirname(path)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod
