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


def _root():
    import os
    return os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_rust_bindings_agree_with_the_header():
    """integration/pathfinder_cuda/src/ffi.rs (source only) uses the header's command kinds, blend / combine /
    filter constants, and declares only entry points the header declares."""
    import os
    import re
    root = _root()
    header = open(os.path.join(root, "include", "pf_cuda.h")).read()
    ffi = open(os.path.join(root, "integration", "pathfinder_cuda", "src", "ffi.rs")).read()
    kinds = {m.group(1): int(m.group(2)) for m in re.finditer(r"PF_RENDER_COMMAND_(\w+) = (\d+)", header)}
    consts = {m.group(1): int(m.group(2), 0) for m in re.finditer(r"pub const (\w+): u32 = (\w+);", ffi)}
    assert len(kinds) == 14
    for name, value in kinds.items():
        assert consts.get(name) == value, name
    # every PF_BLEND_MODE_* / PF_COLOR_COMBINE_MODE_* / PF_FILTER_* of the header has the same value in ffi.rs
    defines = {m.group(1): int(m.group(2), 0) for m in
               re.finditer(r"#define PF_((?:BLEND_MODE|COLOR_COMBINE_MODE|FILTER)_\w+) (\w+)", header)}
    assert len([k for k in defines if k.startswith("BLEND_MODE_")]) == 27
    for name, value in defines.items():
        assert consts.get(name) == value, name
    declared = set(re.findall(r"\b(PF[A-Z]\w+)\s*\(", header))
    for fn in re.findall(r"pub fn (PF\w+)\(", ffi):
        assert fn in declared, fn
    # the extension fields are present on both sides
    for field in ("content_key", "payload_persists", "has_clipped_path_info", "color_texture", "allocate_texture_page",
                  "upload_texel_data", "declare_render_target", "sampling_flags", "composite_op"):
        assert field in header and field in ffi, field
    # texture sampling flags / composite ops (u8 on both sides)
    small = {m.group(1): int(m.group(2), 0) for m in re.finditer(r"pub const (\w+): u8 = (\w+);", ffi)}
    for name, value in re.findall(r"#define PF_((?:TEXTURE_SAMPLING_FLAGS|PAINT_COMPOSITE_OP)_\w+) (\w+)", header):
        assert small.get(name) == int(value, 0), name


def test_rust_bindings_pass_only_verified_pod_records_by_pointer():
    """The glue may hand a `Vec<T>` of a *reference* type to C by pointer only when T is one of the `#[repr(C)]`
    plain-old-data records whose layout pf_cuda.h mirrors (sizes pinned in test_record_layouts and by the `const _`
    assertions in ffi.rs). Everything else — TextureMetadataEntry above all: its Transform2F, Filter and BlendMode
    have no C layout — must be converted into a struct ffi.rs itself declares."""
    import os
    import re
    root = _root()
    ffi = open(os.path.join(root, "integration", "pathfinder_cuda", "src", "ffi.rs")).read()
    lib = open(os.path.join(root, "integration", "pathfinder_cuda", "src", "lib.rs")).read()
    allowed = {"Vector2F", "SegmentIndicesD3D11", "PropagateMetadataD3D11", "DiceMetadataD3D11", "TilePathInfoD3D11",
               "BackdropInfoD3D11", "RectF", "ColorU"}
    own = set(re.findall(r"pub (?:struct|union) (\w+)", ffi))
    primitives = {"u8", "c_char", "c_void", "f32", "u32", "i32", "u64"}
    for type_name in re.findall(r"\*(?:const|mut) (\w+)", ffi):
        assert type_name in allowed | own | primitives, f"ffi.rs passes {type_name} by pointer"
    # each allowed reference record has a size assertion
    for type_name in allowed - {"RectF"}:
        assert re.search(r"size_of::<%s>\(\) == \d+" % type_name, ffi), type_name
    assert "TextureMetadataEntry" not in re.findall(r"\*const (\w+)", ffi)
    # the conversion exists and is used for UploadTextureMetadata
    assert "fn texture_metadata_entry(" in lib and "map(texture_metadata_entry)" in lib
    assert "entries.as_ptr()" not in lib
    # texture pages, render targets and colour textures reach the library with their payloads (no `simple(...)` stub)
    for kind in ("ALLOCATE_TEXTURE_PAGE", "UPLOAD_TEXEL_DATA", "DECLARE_RENDER_TARGET"):
        assert "simple(ffi::%s)" % kind not in lib and "kind: ffi::%s" % kind in lib, kind
    assert "color_texture: tile_batch_texture(" in lib and "fn texture_location(" in lib
    # RendererMode.level is one byte on both sides
    assert re.search(r"pub struct PFRendererMode \{\s*pub level: u8", ffi)


def test_build_rs_compiles_what_the_makefile_compiles():
    """build.rs (source only) and INTEGRATION.md name exactly the sources csrc/Makefile builds — a stale list links
    a library with unresolved symbols."""
    import os
    import re
    root = _root()
    makefile = open(os.path.join(root, "pathfinder_b200", "csrc", "Makefile")).read()
    srcs = set(re.search(r"^SRCS := (.*)$", makefile, re.M).group(1).split())
    assert srcs == {f for f in os.listdir(os.path.join(root, "pathfinder_b200", "csrc")) if f.endswith((".cu", ".cpp"))}
    build_rs = open(os.path.join(root, "integration", "pathfinder_cuda", "build.rs")).read()
    exact = set(re.findall(r'"(\w+\.(?:cu|cpp))"', re.search(r"EXACT_SOURCES[^;]*;", build_rs).group(0)))
    contracted = set(re.findall(r'"(\w+\.(?:cu|cpp))"', re.search(r"CONTRACTED_SOURCES[^;]*;", build_rs).group(0)))
    assert exact | contracted == srcs and not (exact & contracted)
    # the one file the Makefile compiles without -fmad=false is the one build.rs compiles that way
    assert contracted == set(re.findall(r"^\$\(OBJ\)/(\w+)\.o: \w+\.cu", makefile, re.M) and ["composite.cu"])
    assert "$(OBJ)/composite.o: composite.cu" in makefile
    integration_md = open(os.path.join(root, "INTEGRATION.md")).read()
    for f in srcs:
        assert f in integration_md, f"INTEGRATION.md does not mention {f}"


def test_sanitizer_script_builds_the_whole_library():
    """tools/fuzz/run.sh links two instrumented copies of the library (TSan for the scene proxy, ASan for the CPU suite):
    each must name every source csrc/Makefile builds, or the copy fails to load and the step drops out unnoticed."""
    import os
    import re
    root = _root()
    makefile = open(os.path.join(root, "pathfinder_b200", "csrc", "Makefile")).read()
    srcs = set(re.search(r"^SRCS := (.*)$", makefile, re.M).group(1).split())
    script = open(os.path.join(root, "tools", "fuzz", "run.sh")).read().replace("\\\n", " ")
    whole = [line for line in script.splitlines() if "kernels.cu" in line]
    assert len(whole) == 2
    for line in whole:
        assert set(re.findall(r"(\w+\.(?:cu|cpp))\b", line)) - {"proxy_tsan.cpp"} == srcs, line
