"""The C-ABI library loads on a CPU-only box and exports every symbol include/pf_cuda.h declares
(no compute calls here)."""
import ctypes
import os
import re

from pathfinder_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "pf_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(PF[A-Za-z0-9]+)\s*\(", text)
    # `typedef PFCudaStatus (*PFRenderCommandListenerFn)(...)` is a type, not an entry point
    return sorted({n for n in names if not n.endswith("Fn") and n != "PFCudaStatus"})


def test_header_declares_entry_points():
    names = declared_functions()
    for must in ("PFCudaRendererCreate", "PFCudaRendererBeginScene", "PFCudaRendererRenderCommand",
                 "PFCudaRendererEndScene", "PFSceneBuild", "PFSceneBuildAndRenderCuda"):
        assert must in names


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, f"libpf_cuda.so lacks {missing}"


def test_bindings_cover_the_header():
    assert sorted(_lib.SIGNATURES) == declared_functions()
    _lib.lib()


def test_no_device_is_a_loud_error():
    """Without a GPU the product path fails (no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    lib = _lib.lib()
    assert not lib.PFCudaDeviceCreate(0)
    assert b"CUDA" in lib.PFCudaGetLastError()


def test_record_layouts():
    # gpu_data.rs #[repr(C)] sizes
    assert ctypes.sizeof(_lib.PFSegmentIndicesD3D11) == 8
    assert ctypes.sizeof(_lib.PFDiceMetadataD3D11) == 16
    assert ctypes.sizeof(_lib.PFTilePathInfoD3D11) == 16
    assert ctypes.sizeof(_lib.PFPropagateMetadataD3D11) == 48
    assert ctypes.sizeof(_lib.PFBackdropInfoD3D11) == 12
    from pathfinder_b200 import api
    assert api.FILL_DTYPE.itemsize == 12 and api.TILE_DTYPE.itemsize == 16


def test_rust_bindings_agree_with_the_header():
    """integration/pathfinder_cuda/src/ffi.rs (source only) uses the header's command kinds and declares only
    entry points the header declares."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "pf_cuda.h")).read()
    ffi = open(os.path.join(root, "integration", "pathfinder_cuda", "src", "ffi.rs")).read()
    kinds = {m.group(1): int(m.group(2)) for m in re.finditer(r"PF_RENDER_COMMAND_(\w+) = (\d+)", header)}
    consts = {m.group(1): int(m.group(2)) for m in re.finditer(r"pub const (\w+): u32 = (\d+);", ffi)}
    assert len(kinds) == 14 and consts == kinds
    declared = set(re.findall(r"\b(PF[A-Z]\w+)\s*\(", header))
    for fn in re.findall(r"pub fn (PF\w+)\(", ffi):
        assert fn in declared, fn
    # the extension fields are present on both sides
    for field in ("content_key", "payload_persists", "has_clipped_path_info"):
        assert field in header and field in ffi
