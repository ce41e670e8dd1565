"""pathfinder_b200 — a B200-native `cuda` backend for Pathfinder 3's D3D11-level rasterization
pipeline (bound -> dice -> bin -> propagate -> sort -> fill -> tile), behind the
`pathfinder_renderer` API surface. See DESIGN.md and include/pf_cuda.h.

The package holds the CUDA kernels and C ABI (csrc/ -> libpf_cuda.so), the Python mirror of the
reference's Scene / Renderer interface (api.py), the area-LUT generator and scene generators for
the benchmark configurations. There is no CPU implementation of the pipeline in here."""
from .flat_scene import FILL_RULE_EVEN_ODD, FILL_RULE_WINDING, FlatScene, SceneBuilderPy  # noqa: F401


def __getattr__(name):
    # api pulls in ctypes bindings lazily so that scene generation works without the .so.
    if name in ("Scene", "BuildOptions", "CudaRenderer", "Transform2F", "RendererLevel"):
        from . import api
        return getattr(api, name)
    raise AttributeError(name)
