"""ctypes bindings for the CPU checker under oracle/ (test infrastructure only).

`Oracle` wraps oracle/libquasi_oracle.so, the plain-C restatement of the
reference shader (/root/reference/backends/metal/plugin.mm:98-350).
`Reference` wraps oracle/_ref/libquasi_ref.so, the reference's own shader text
and scene header compiled from /root/reference (present only where that tree
was available at build time). Only tests, smoke() and bench.py's CPU legs may
import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
ORACLE_SO = ORACLE_DIR / "libquasi_oracle.so"
REF_SO = ORACLE_DIR / "_ref" / "libquasi_ref.so"
REFERENCE_TREE = Path("/root/reference")

UNIFORMS_BYTES = 2624  # sizeof(GpuSceneUniforms), plugin.mm:379-387
MAX_QUADS = 32

# Default host camera: orbit_camera in /root/reference/src/quasi/host/main.cpp:37-45
DEFAULT_CAMERA = dict(position=(0.0, 1.0, 3.5), target=(0.0, 1.0, 0.0), up=(0.0, 1.0, 0.0), fov=40.0)

uniforms_dtype = np.dtype(
    [
        ("cam_position", "<f4", 3),
        ("cam_fov", "<f4"),
        ("cam_direction", "<f4", 3),
        ("cam_aspect", "<f4"),
        ("cam_up", "<f4", 3),
        ("cam_pad", "<f4"),
        ("quad_count", "<u4"),
        ("frame_count", "<u4"),
        ("light_index", "<u4"),
        ("pad", "<f4"),
        ("quads", "<f4", (MAX_QUADS, 12)),
        ("materials", "<f4", (MAX_QUADS, 8)),
    ]
)
assert uniforms_dtype.itemsize == UNIFORMS_BYTES


def build_oracle(ref: bool = True) -> None:
    """Compile the checker (idempotent; make decides what is stale)."""
    subprocess.run(["make", "-s", "-C", str(ORACLE_DIR), "all"], check=True)
    if ref and REFERENCE_TREE.exists():
        subprocess.run(["make", "-s", "-C", str(ORACLE_DIR), "ref"], check=True)


def _fptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


class _Base:
    def new_uniforms(self) -> np.ndarray:
        return np.zeros((), dtype=uniforms_dtype)

    def alloc_layers(self, width: int, height: int):
        return [np.zeros((height, width, 4), dtype=np.float32) for _ in range(4)]


class Oracle(_Base):
    def __init__(self, path: Path = ORACLE_SO):
        if not path.exists():
            build_oracle(ref=False)
        self.lib = lib = C.CDLL(str(path))
        u32, u64, f32, vp, i32 = C.c_uint32, C.c_uint64, C.c_float, C.c_void_p, C.c_int
        fp = C.POINTER(C.c_float)
        lib.qo_pcg_hash.restype = u32
        lib.qo_pcg_hash.argtypes = [u32]
        lib.qo_random_float.restype = f32
        lib.qo_random_float.argtypes = [C.POINTER(u32)]
        lib.qo_seed.restype = u32
        lib.qo_seed.argtypes = [u32, u32, u32]
        lib.qo_make_cornell_box.restype = None
        lib.qo_make_cornell_box.argtypes = [vp]
        lib.qo_pack_camera.restype = None
        lib.qo_pack_camera.argtypes = [vp, fp, fp, fp, f32, u32, u32]
        lib.qo_pathtrace_pixel.restype = u32
        lib.qo_pathtrace_pixel.argtypes = [vp, u32, u32, u32, u32, fp]
        lib.qo_render_sample.restype = u64
        lib.qo_render_sample.argtypes = [vp, u32, u32, fp, fp, fp, fp, i32]
        lib.qo_accumulate.restype = None
        lib.qo_accumulate.argtypes = [fp, fp, fp, u64, u32]
        lib.qo_render_frames.restype = u64
        lib.qo_render_frames.argtypes = [vp, u32, u32, u32, u32, u32, C.POINTER(fp), i32]
        lib.qo_tonemap_bgra8.restype = None
        lib.qo_tonemap_bgra8.argtypes = [fp, u32, u32, C.POINTER(C.c_uint8)]
        lib.qo_combine_shards.restype = None
        lib.qo_combine_shards.argtypes = [C.POINTER(fp), u32, u64, fp]
        for n in ("qo_sin", "qo_cos", "qo_tan"):
            getattr(lib, n).restype = f32
            getattr(lib, n).argtypes = [f32]
        lib.qo_sincos_array.restype = None
        lib.qo_sincos_array.argtypes = [fp, u64, fp, fp]
        lib.qo_online_cores.restype = i32

    # integer path
    def pcg_hash(self, x: int) -> int:
        return self.lib.qo_pcg_hash(x & 0xFFFFFFFF)

    def seed(self, px: int, py: int, frame: int) -> int:
        return self.lib.qo_seed(px, py, frame)

    def random_floats(self, state: int, n: int):
        s = C.c_uint32(state)
        states, floats = [], []
        for _ in range(n):
            f = self.lib.qo_random_float(C.byref(s))
            states.append(s.value)
            floats.append(np.float32(f))
        return states, floats

    def cores(self) -> int:
        return self.lib.qo_online_cores()

    # scene
    def scene(self, width: int, height: int, camera: dict = DEFAULT_CAMERA, frame: int = 0):
        u = self.new_uniforms()
        self.lib.qo_make_cornell_box(u.ctypes.data)
        self.lib.qo_pack_camera(
            u.ctypes.data, _f3(camera["position"]), _f3(camera["target"]), _f3(camera["up"]),
            float(camera["fov"]), width, height)
        u["frame_count"] = frame
        return u

    def pathtrace_pixel(self, u, px, py, width, height):
        out = np.zeros(16, dtype=np.float32)
        rays = self.lib.qo_pathtrace_pixel(u.ctypes.data, px, py, width, height, _fptr(out))
        return out.reshape(4, 4), rays

    def render_sample(self, u, width, height, threads=0):
        layers = self.alloc_layers(width, height)
        rays = self.lib.qo_render_sample(u.ctypes.data, width, height, *[_fptr(a) for a in layers], threads)
        return layers, rays

    def accumulate(self, current, history, frame_count):
        out = np.empty_like(history)
        self.lib.qo_accumulate(_fptr(current), _fptr(history), _fptr(out), current.size // 4, frame_count)
        return out

    def render_frames(self, u, width, height, n_frames, frame_first=0, frame_stride=1, threads=0):
        layers = self.alloc_layers(width, height)
        arr = (C.POINTER(C.c_float) * 4)(*[_fptr(a) for a in layers])
        rays = self.lib.qo_render_frames(u.ctypes.data, width, height, frame_first, frame_stride,
                                         n_frames, arr, threads)
        return layers, rays

    def tonemap_bgra8(self, accumulated):
        h, w, _ = accumulated.shape
        out = np.zeros((h, w, 4), dtype=np.uint8)
        self.lib.qo_tonemap_bgra8(_fptr(np.ascontiguousarray(accumulated)), w, h,
                                  out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out

    def sincos(self, phi):
        phi = np.ascontiguousarray(phi, dtype=np.float32)
        s, c = np.empty_like(phi), np.empty_like(phi)
        self.lib.qo_sincos_array(_fptr(phi), phi.size, _fptr(s), _fptr(c))
        return s, c

    def combine_shards(self, shards):
        shards = [np.ascontiguousarray(s, dtype=np.float32) for s in shards]
        out = np.empty_like(shards[0])
        arr = (C.POINTER(C.c_float) * len(shards))(*[_fptr(s) for s in shards])
        self.lib.qo_combine_shards(arr, len(shards), out.size, _fptr(out))
        return out


class Reference(_Base):
    """The reference's own shader text + scene header (oracle/_ref)."""

    @staticmethod
    def available() -> bool:
        return REF_SO.exists()

    def __init__(self, path: Path = REF_SO):
        self.lib = lib = C.CDLL(str(path))
        u32, f32, vp, i32 = C.c_uint32, C.c_float, C.c_void_p, C.c_int
        fp = C.POINTER(C.c_float)
        lib.ref_msl_sizeof_uniforms.restype = u32
        lib.ref_msl_pcg_hash.restype = u32
        lib.ref_msl_pcg_hash.argtypes = [u32]
        lib.ref_msl_random_float.restype = f32
        lib.ref_msl_random_float.argtypes = [C.POINTER(u32)]
        lib.ref_msl_pathtrace_frag.restype = None
        lib.ref_msl_pathtrace_frag.argtypes = [vp, u32, u32, u32, u32, fp]
        lib.ref_make_cornell_box.restype = None
        lib.ref_make_cornell_box.argtypes = [f32, vp]
        lib.ref_cornell_box_camera.restype = None
        lib.ref_cornell_box_camera.argtypes = [f32, fp]
        lib.ref_pack_camera.restype = None
        lib.ref_pack_camera.argtypes = [vp, fp, fp, fp, f32, u32, u32]
        lib.ref_render_sample.restype = None
        lib.ref_render_sample.argtypes = [vp, u32, u32, fp, fp, fp, fp, i32]
        lib.ref_accumulate.restype = None
        lib.ref_accumulate.argtypes = [fp, fp, fp, u32, u32, u32, i32]
        lib.ref_render_frames.restype = None
        lib.ref_render_frames.argtypes = [vp, u32, u32, u32, u32, u32, C.POINTER(fp), i32]
        lib.ref_tonemap_bgra8.restype = None
        lib.ref_tonemap_bgra8.argtypes = [fp, u32, u32, C.POINTER(C.c_uint8), i32]
        lib.ref_hardware_threads.restype = i32
        assert lib.ref_msl_sizeof_uniforms() == UNIFORMS_BYTES

    def pcg_hash(self, x: int) -> int:
        return self.lib.ref_msl_pcg_hash(x & 0xFFFFFFFF)

    def random_floats(self, state: int, n: int):
        s = C.c_uint32(state)
        states, floats = [], []
        for _ in range(n):
            f = self.lib.ref_msl_random_float(C.byref(s))
            states.append(s.value)
            floats.append(np.float32(f))
        return states, floats

    def cores(self) -> int:
        return self.lib.ref_hardware_threads()

    def scene(self, width: int, height: int, camera: dict = DEFAULT_CAMERA, frame: int = 0):
        u = self.new_uniforms()
        # plugin.mm:582-584: aspect = viewport_width / viewport_height
        aspect = float(np.float32(width) / np.float32(height))
        self.lib.ref_make_cornell_box(aspect, u.ctypes.data)
        self.lib.ref_pack_camera(
            u.ctypes.data, _f3(camera["position"]), _f3(camera["target"]), _f3(camera["up"]),
            float(camera["fov"]), width, height)
        u["frame_count"] = frame
        return u

    def scene_camera(self, aspect: float):
        out = np.zeros(11, dtype=np.float32)
        self.lib.ref_cornell_box_camera(aspect, _fptr(out))
        return out

    def pathtrace_pixel(self, u, px, py, width, height):
        out = np.zeros(16, dtype=np.float32)
        self.lib.ref_msl_pathtrace_frag(u.ctypes.data, px, py, width, height, _fptr(out))
        return out.reshape(4, 4)

    def render_sample(self, u, width, height, threads=0):
        layers = self.alloc_layers(width, height)
        self.lib.ref_render_sample(u.ctypes.data, width, height, *[_fptr(a) for a in layers], threads)
        return layers

    def accumulate(self, current, history, frame_count, threads=0):
        h, w, _ = current.shape
        out = np.empty_like(history)
        self.lib.ref_accumulate(_fptr(current), _fptr(history), _fptr(out), w, h, frame_count, threads)
        return out

    def render_frames(self, u, width, height, n_frames, frame_first=0, frame_stride=1, threads=0):
        layers = self.alloc_layers(width, height)
        arr = (C.POINTER(C.c_float) * 4)(*[_fptr(a) for a in layers])
        self.lib.ref_render_frames(u.ctypes.data, width, height, frame_first, frame_stride,
                                   n_frames, arr, threads)
        return layers

    def tonemap_bgra8(self, accumulated, threads=0):
        h, w, _ = accumulated.shape
        out = np.zeros((h, w, 4), dtype=np.uint8)
        self.lib.ref_tonemap_bgra8(_fptr(np.ascontiguousarray(accumulated)), w, h,
                                   out.ctypes.data_as(C.POINTER(C.c_uint8)), threads)
        return out


def bits(a: np.ndarray) -> np.ndarray:
    """View float32 data as uint32 for bit-exact comparison."""
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


class MeshOracle:
    """oracle/quasi_oracle_mesh.c: the triangle-mesh scene of BASELINE config 3."""

    node_dtype = np.dtype([("bmin", "<f4", 3), ("left_first", "<u4"), ("bmax", "<f4", 3), ("count", "<u4")])
    tri_dtype = np.dtype([("v0", "<f4", 3), ("material", "<u4"), ("e1", "<f4", 3), ("pad0", "<u4"),
                          ("e2", "<f4", 3), ("pad1", "<u4")])

    class _stats(C.Structure):
        _fields_ = [("rays", C.c_uint64), ("nodes_visited", C.c_uint64), ("triangles_tested", C.c_uint64)]

    def __init__(self, grid: int, seed: int = 1, path: Path = ORACLE_SO):
        if not path.exists():
            build_oracle(ref=False)
        self.lib = lib = C.CDLL(str(path))
        fp = C.POINTER(C.c_float)
        lib.qm_scene_create.restype = C.c_void_p
        lib.qm_scene_create.argtypes = [C.c_uint32, C.c_uint32]
        lib.qm_scene_destroy.argtypes = [C.c_void_p]
        for n in ("qm_triangle_count", "qm_node_count"):
            getattr(lib, n).restype = C.c_uint32
            getattr(lib, n).argtypes = [C.c_void_p]
        for n in ("qm_triangles", "qm_nodes", "qm_leaf_order"):
            getattr(lib, n).restype = C.c_void_p
            getattr(lib, n).argtypes = [C.c_void_p]
        lib.qm_render_frames.restype = None
        lib.qm_render_frames.argtypes = [C.c_void_p, fp, fp, fp, C.c_float, C.c_uint32, C.c_uint32, C.c_uint32,
                                         C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(fp), C.c_int,
                                         C.POINTER(self._stats)]
        self.grid, self.seed = grid, seed
        self.scene = lib.qm_scene_create(grid, seed)

    def __del__(self):
        try:
            self.lib.qm_scene_destroy(self.scene)
        except Exception:
            pass

    @property
    def n_triangles(self):
        return self.lib.qm_triangle_count(self.scene)

    @property
    def n_nodes(self):
        return self.lib.qm_node_count(self.scene)

    def _array(self, ptr, dtype, n):
        buf = (C.c_char * (np.dtype(dtype).itemsize * n)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype, count=n).copy()

    def triangles(self):
        return self._array(self.lib.qm_triangles(self.scene), self.tri_dtype, self.n_triangles)

    def nodes(self):
        return self._array(self.lib.qm_nodes(self.scene), self.node_dtype, self.n_nodes)

    def leaf_order(self):
        return self._array(self.lib.qm_leaf_order(self.scene), np.uint32, self.n_triangles)

    def render_frames(self, width, height, n_frames, max_bounces=4, use_bvh=True, camera=DEFAULT_CAMERA,
                      frame_first=0, frame_stride=1, threads=0):
        layers = [np.zeros((height, width, 4), dtype=np.float32) for _ in range(4)]
        arr = (C.POINTER(C.c_float) * 4)(*[_fptr(a) for a in layers])
        st = self._stats()
        self.lib.qm_render_frames(self.scene, _f3(camera["position"]), _f3(camera["target"]), _f3(camera["up"]),
                                  float(camera["fov"]), width, height, frame_first, frame_stride, n_frames,
                                  max_bounces, 1 if use_bvh else 0, arr, threads, C.byref(st))
        return layers, dict(rays=st.rays, nodes_visited=st.nodes_visited, triangles_tested=st.triangles_tested)
