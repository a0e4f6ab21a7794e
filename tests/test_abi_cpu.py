"""CPU-only checks of the drop-in boundary: the library loads, exports every
symbol include/*.h declares, and the struct layouts match the reference ABI.
No compute entry point is called here (there is no GPU in this container)."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

import quasi_b200 as q
from quasi_b200 import abi
from quasi_b200.loader import plugin_error

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def library():
    if not q.PLUGIN_PATH.exists():
        subprocess.run(["make", "-s", "-C", str(ROOT), "plugin"], check=True)
    return q.dynamic_library.open(q.PLUGIN_PATH)


def declared_symbols(header: Path):
    text = header.read_text()
    return re.findall(r"Q_PLUGIN_EXPORT[^;(]*?\b(Q_\w+)\s*\(", text)


def test_library_exports_every_declared_symbol(library):
    names = declared_symbols(ROOT / "include/quasi_plugin_abi.h") + declared_symbols(
        ROOT / "include/quasi_cuda_backend.h")
    assert len(names) == 32 and set(abi.REQUIRED_SYMBOLS + abi.OPTIONAL_SYMBOLS) <= set(names)
    for n in names:
        assert library.has_symbol(n), n


def test_abi_version_and_info_need_no_gpu(library):
    v = library.get_symbol("Q_plugin_abi_version")
    v.restype = C.c_uint32
    assert 1 <= v() <= abi.K_PLUGIN_ABI_VERSION and v() == 3
    info = library.get_symbol("Q_plugin_get_info")
    info.restype = abi.Q_plugin_info
    i = info()
    assert i.name == b"CUDA Path Tracer" and (i.version.major, i.version.minor, i.version.patch) == (1, 0, 0)


def test_null_tolerance_without_gpu(library):
    # plugin.mm:557, 595, 625, 848, 866, 874, 895: NULL arguments are silently ignored.
    create = library.get_symbol("Q_plugin_create")
    create.restype = C.c_void_p
    create.argtypes = [C.c_void_p]
    assert create(None) is None
    ctx = abi.Q_plugin_context()  # gpu == NULL
    assert create(C.addressof(ctx)) is None
    for name in ("Q_plugin_destroy", "Q_plugin_readback_free", "Q_plugin_readback_aov_free"):
        f = library.get_symbol(name)
        f.restype = None
        f.argtypes = [C.c_void_p]
        f(None)
    r = library.get_symbol("Q_plugin_render")
    r.restype = None
    r.argtypes = [C.c_void_p, C.c_void_p]
    r(None, None)
    rb = library.get_symbol("Q_plugin_readback")
    rb.restype = abi.Q_readback_result
    rb.argtypes = [C.c_void_p]
    assert not rb(None).data
    rba = library.get_symbol("Q_plugin_readback_aov")
    rba.restype = abi.Q_readback_aov_result
    rba.argtypes = [C.c_void_p]
    assert all(not b.data for b in rba(None).buffers)


def test_wrong_backend_is_rejected_with_a_log_line(library):
    # plugin.mm:561-566
    messages = []
    gpu = abi.Q_gpu_context(abi.Q_GPU_BACKEND_METAL, None, None, None)
    ctx = abi.Q_plugin_context()
    ctx.viewport_width = ctx.viewport_height = 8
    ctx.gpu = C.pointer(gpu)
    cb = abi.LOG_FN(lambda _h, m: messages.append(m.decode()))
    ctx.log = cb
    with pytest.raises(plugin_error) as e:
        q.loader.load(library, ctx)
    assert e.value.code is q.loader_error.create_failed
    assert messages == ["CUDA backend requires CUDA GPU context"]


def test_loader_error_cases():
    with pytest.raises(plugin_error) as e:
        q.dynamic_library.open("/nonexistent/libnothing.so")
    assert e.value.code is q.library_error.file_not_found
    with pytest.raises(plugin_error) as e:
        q.dynamic_library().get_symbol("x")
    assert e.value.code is q.library_error.not_loaded
    libm = q.dynamic_library.open(Path("/lib/x86_64-linux-gnu/libm.so.6"))
    with pytest.raises(plugin_error) as e:
        q.loader.load(libm, abi.Q_plugin_context())
    assert e.value.code is q.loader_error.symbol_not_found


def test_struct_layouts_match_the_c_headers(tmp_path):
    # Compile a C program against include/*.h and compare sizeof/offsetof with ctypes.
    src = tmp_path / "layout.c"
    src.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "quasi_cuda_backend.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(Q_plugin_context), sizeof(Q_render_frame),
    sizeof(Q_gpu_context), sizeof(Q_camera), sizeof(Q_readback_result), sizeof(Q_readback_aov_result),
    sizeof(Q_cuda_device_desc), sizeof(Q_cuda_drawable), sizeof(Q_cuda_stats), sizeof(Q_cuda_layers));
  printf("%zu %zu %zu %zu\n", sizeof(Q_cuda_persistent_state), offsetof(Q_cuda_persistent_state, layers),
    offsetof(Q_cuda_persistent_state, frame_count), offsetof(Q_cuda_persistent_state, device_ordinal));
  printf("%zu %zu %zu %zu\n", offsetof(Q_render_frame, camera), offsetof(Q_render_frame, camera_dirty),
    offsetof(Q_plugin_context, gpu), offsetof(Q_plugin_context, request_shutdown));
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(Q_cuda_export_table), offsetof(Q_cuda_export_table, base),
    sizeof(Q_cuda_reduce_table), offsetof(Q_cuda_reduce_table, staging), offsetof(Q_cuda_reduce_table, counters),
    offsetof(Q_cuda_reduce_table, out));
  return 0; }''')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", f"-I{ROOT / 'include'}", str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    sizes = [C.sizeof(t) for t in (abi.Q_plugin_context, abi.Q_render_frame, abi.Q_gpu_context, abi.Q_camera,
                                   abi.Q_readback_result, abi.Q_readback_aov_result, abi.Q_cuda_device_desc,
                                   abi.Q_cuda_drawable, abi.Q_cuda_stats, abi.Q_cuda_layers)]
    assert [int(x) for x in out[:10]] == sizes
    assert sizes[:6] == [40, 72, 32, 40, 24, 96]  # SURVEY.md 8b, probed from the reference headers
    P = abi.Q_cuda_persistent_state
    assert [int(x) for x in out[10:14]] == [C.sizeof(P), P.layers.offset, P.frame_count.offset, P.device_ordinal.offset]
    assert [int(x) for x in out[14:18]] == [abi.Q_render_frame.camera.offset, abi.Q_render_frame.camera_dirty.offset,
                                            abi.Q_plugin_context.gpu.offset, abi.Q_plugin_context.request_shutdown.offset]
    E, R = abi.Q_cuda_export_table, abi.Q_cuda_reduce_table  # the multi-GPU tables of quasi_cuda_backend.h
    assert [int(x) for x in out[18:]] == [C.sizeof(E), E.base.offset, C.sizeof(R), R.staging.offset, R.counters.offset,
                                          R.out.offset]


@pytest.mark.skipif(not Path("/root/reference/src").exists(), reason="reference tree not present")
def test_struct_layouts_match_the_reference_headers(tmp_path):
    # The reference's own headers, compiled from where they lie, must agree with ours.
    src = tmp_path / "layout.cpp"
    src.write_text(r'''
#include <quasi/plugin/plugin_interface.hpp>
namespace ours {
#include "quasi_plugin_abi.h"
}
#include <cstddef>
#define SAME(T) static_assert(sizeof(::T) == sizeof(ours::T) && alignof(::T) == alignof(ours::T), #T)
SAME(Q_plugin_context); SAME(Q_render_frame); SAME(Q_gpu_context); SAME(Q_camera); SAME(Q_plugin_info);
SAME(Q_readback_result); SAME(Q_aov_buffer); SAME(Q_readback_aov_result); SAME(Q_plugin_version);
static_assert(offsetof(::Q_render_frame, camera_dirty) == offsetof(ours::Q_render_frame, camera_dirty));
static_assert(offsetof(::Q_plugin_context, request_shutdown) == offsetof(ours::Q_plugin_context, request_shutdown));
static_assert(offsetof(::Q_gpu_context, layer) == offsetof(ours::Q_gpu_context, layer));
static_assert(sizeof(::Q_gpu_backend) == sizeof(ours::Q_gpu_backend) && sizeof(::Q_aov_type) == sizeof(ours::Q_aov_type));
static_assert(Q::plugin::k_plugin_abi_version == Q_PLUGIN_ABI_VERSION);
static_assert(int(::Q_AOV_DEPTH) == int(ours::Q_AOV_DEPTH) && int(ours::Q_GPU_BACKEND_CUDA) == 4);
int main() { return 0; }''')
    subprocess.run(["g++", "-std=c++20", "-fsyntax-only", "-I/root/reference/src", f"-I{ROOT / 'include'}", str(src)],
                   check=True)


def test_every_environment_option_of_the_plugin_is_documented():
    """The ABI has no slot for options, so the plugin reads QUASI_CUDA_* from the environment
    (SURVEY.md 8b); INTEGRATION.md is where a maintainer finds them."""
    import re
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    read = set(re.findall(r'"(QUASI_CUDA_[A-Z_]+)"', (root / "quasi_b200" / "csrc" / "plugin.cu").read_text()))
    doc = (root / "INTEGRATION.md").read_text()
    assert read and not [name for name in sorted(read) if name not in doc]
