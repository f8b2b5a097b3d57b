"""petershirleyraytracer_b200 -- B200-native (sm_100a) path-tracing hot path of fengye/PeterShirleyRaytracer.

This package is a thin ctypes binding of the C ABI in include/rt.h (librt_b200.so, hand-written CUDA).
It is plumbing for tests, bench.py and the multi-GPU driver; the product is the CUDA library and the C++
host API in include/.  There is no CPU implementation here: if the library is missing, importing
`lib()` raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RT_B200_LIB") or os.path.join(PKG_DIR, "librt_b200.so")  # env override: tuning builds only

SCAN_FILTERED, SCAN_EXACT, SCAN_BVH, SCAN_AUTO = 0, 1, 2, 3
SCATTER_HEMISPHERE, SCATTER_LAMBERTIAN = 0, 1
ABI_VERSION = 3
TILE_W, TILE_H = 8, 8

# every symbol include/rt.h declares (tests check the built library exports all of them)
ABI_SYMBOLS = [
    "rt_abi_version", "rt_last_error", "rt_upload_scene", "rt_free_scene", "rt_scene_size", "rt_render",
    "rt_render_device", "rt_render_finish", "rt_get_tile_layout", "rt_deinterleave", "rt_primary_hits", "rt_hit",
    "rt_ray_color", "rt_write_color", "rt_get_ray", "rt_philox", "rt_measure_fp32_peak", "rt_device_info",
    "rt_check_division", "rt_update_scene", "rt_accum_bytes", "rt_render_pass", "rt_render_pass_device",
    "rt_params_init", "rt_ray_color_params", "rt_accum_create", "rt_accum_destroy", "rt_accum_samples", "rt_accum_reset",
    "rt_accum_add", "rt_accum_frame", "rt_accum_read", "rt_accum_write", "rt_accum_to_frame", "rt_render_multi",
]


class RtCamera(C.Structure):
    _fields_ = [("origin", C.c_double * 3), ("lower_left_corner", C.c_double * 3),
                ("horizontal", C.c_double * 3), ("vertical", C.c_double * 3)]


class RtParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("spp", C.c_int32), ("max_depth", C.c_int32),
                ("seed", C.c_uint64), ("tmin", C.c_double), ("jitter", C.c_int32), ("early_out", C.c_int32),
                ("scan_mode", C.c_int32), ("shard_rank", C.c_int32), ("shard_count", C.c_int32),
                ("reserved", C.c_int32 * 3),
                ("custom_shading", C.c_int32), ("scatter_mode", C.c_int32), ("albedo", C.c_double),
                ("sky_a", C.c_double * 3), ("sky_b", C.c_double * 3)]


class RtStats(C.Structure):
    _fields_ = [("kernel_ms", C.c_double)] + [(n, C.c_uint64) for n in (
        "samples", "casts", "sphere_tests", "node_tests", "exact_tests", "black", "early_outs", "primary_hits",
        "overflows", "launches", "self_resolved")]

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_}


class RtTileLayout(C.Structure):
    _fields_ = [("tile_w", C.c_int32), ("tile_h", C.c_int32), ("tiles_x", C.c_int32), ("tiles_y", C.c_int32),
                ("tiles_total", C.c_int32), ("tiles_per_shard", C.c_int32), ("shard_bytes", C.c_int64)]


class RtError(RuntimeError):
    pass


_lib = None


def lib() -> C.CDLL:
    """Loads librt_b200.so (fails loudly: there is no fallback path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RtError(f"{LIB_PATH} not built: run `python -m petershirleyraytracer_b200.build` "
                      "(or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_void_p
    L.rt_abi_version.restype = C.c_int
    L.rt_last_error.restype = C.c_char_p
    L.rt_upload_scene.argtypes = [dp, dp, C.c_int32, C.c_int32, C.POINTER(vp)]
    L.rt_free_scene.argtypes = [vp]
    L.rt_free_scene.restype = None
    L.rt_scene_size.argtypes = [vp]
    L.rt_render.argtypes = [vp, C.POINTER(RtCamera), C.POINTER(RtParams), C.POINTER(C.c_uint8), dp, C.POINTER(RtStats)]
    L.rt_render_device.argtypes = [vp, C.POINTER(RtCamera), C.POINTER(RtParams), vp, vp, vp]
    L.rt_render_finish.argtypes = [vp, C.POINTER(RtStats)]
    L.rt_get_tile_layout.argtypes = [C.POINTER(RtParams), C.POINTER(RtTileLayout)]
    L.rt_deinterleave.argtypes = [C.POINTER(RtParams), vp, vp, C.c_int32, vp]
    L.rt_primary_hits.argtypes = [vp, C.POINTER(RtCamera), C.POINTER(RtParams), ip, dp]
    L.rt_hit.argtypes = [vp, dp, dp, C.c_int32, C.c_double, C.c_double, C.c_int32, ip, dp]
    L.rt_ray_color.argtypes = [vp, dp, dp, C.c_int32, C.c_int32, C.c_uint64, C.c_int32, C.c_int32, dp, C.POINTER(RtStats)]
    L.rt_write_color.argtypes = [dp, C.c_int32, C.c_int32, C.c_int32, ip]
    L.rt_get_ray.argtypes = [C.POINTER(RtCamera), dp, C.c_int32, C.c_int32, dp]
    L.rt_philox.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_int32, C.c_int32, C.POINTER(C.c_uint32)]
    L.rt_measure_fp32_peak.argtypes = [C.c_int32, dp, dp]
    L.rt_device_info.argtypes = [C.c_int32, ip, ip, ip, C.c_char_p, C.c_int32]
    L.rt_update_scene.argtypes = [C.c_void_p, dp, dp, C.c_int32, C.c_int32]
    L.rt_accum_bytes.restype = C.c_int64
    L.rt_accum_bytes.argtypes = [C.POINTER(RtParams)]
    L.rt_render_pass.argtypes = [C.c_void_p, C.POINTER(RtCamera), C.POINTER(RtParams), C.c_int32, C.POINTER(C.c_uint64),
                                 C.POINTER(C.c_uint8), C.POINTER(RtStats)]
    L.rt_render_pass_device.argtypes = [C.c_void_p, C.POINTER(RtCamera), C.POINTER(RtParams), C.c_int32, C.c_void_p,
                                        C.c_void_p, C.c_void_p]
    L.rt_check_division.argtypes = [C.c_int32, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
    L.rt_params_init.argtypes = [C.POINTER(RtParams), C.c_int32, C.c_int32, C.c_int32, C.c_int32]
    L.rt_ray_color_params.argtypes = [vp, dp, dp, C.c_int32, C.POINTER(RtParams), dp, C.POINTER(RtStats)]
    L.rt_accum_create.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.POINTER(vp)]
    L.rt_accum_destroy.argtypes = [vp]
    L.rt_accum_destroy.restype = None
    L.rt_accum_samples.argtypes = [vp]
    L.rt_accum_reset.argtypes = [vp]
    L.rt_accum_add.argtypes = [vp, C.POINTER(RtCamera), C.POINTER(RtParams), vp, C.POINTER(RtStats)]
    L.rt_accum_frame.argtypes = [vp, C.POINTER(C.c_uint8)]
    L.rt_accum_read.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.rt_accum_write.argtypes = [vp, C.POINTER(C.c_uint64), C.c_int32]
    L.rt_accum_to_frame.argtypes = [C.POINTER(RtParams), vp, C.c_int32, vp, C.c_int32, vp]
    L.rt_render_multi.argtypes = [C.POINTER(vp), C.c_int32, C.POINTER(RtCamera), C.POINTER(RtParams), C.POINTER(C.c_uint8),
                                  C.POINTER(RtStats)]
    if L.rt_abi_version() != ABI_VERSION:
        raise RtError(f"{LIB_PATH} has ABI {L.rt_abi_version()}, this binding needs {ABI_VERSION}: rebuild the library")
    _lib = L
    return L


def _check(rc: int) -> None:
    if rc != 0:
        raise RtError(f"rt error {rc}: {lib().rt_last_error().decode(errors='replace')}")


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a, shape=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


@dataclass
class Camera:
    """The public fields of the reference camera (programs/camera.h:31-35) that get_ray reads."""
    origin: np.ndarray
    lower_left_corner: np.ndarray
    horizontal: np.ndarray
    vertical: np.ndarray

    @staticmethod
    def default() -> "Camera":
        """camera::camera() of programs/camera.h:11-23 (same operation order)."""
        aspect = 16.0 / 9.0
        vh = 2.0
        vw = vh * aspect
        origin = np.zeros(3)
        hor = np.array([vw, 0.0, 0.0])
        ver = np.array([0.0, vh, 0.0])
        llc = origin - (1 / 2.0) * hor - (1 / 2.0) * ver + np.array([0.0, 0.0, -1.0])
        return Camera(origin, llc, hor, ver)

    @staticmethod
    def look_at(lookfrom, lookat, vup, vfov_deg: float, aspect: float) -> "Camera":
        """Pinhole camera expressed through the reference's public fields."""
        lookfrom, lookat, vup = _f64(lookfrom), _f64(lookat), _f64(vup)
        w = lookfrom - lookat
        w = w / np.sqrt(w @ w)
        u = np.cross(vup, w)
        u = u / np.sqrt(u @ u)
        v = np.cross(w, u)
        vh = 2.0 * np.tan(np.radians(vfov_deg) / 2.0)
        vw = vh * aspect
        hor, ver = vw * u, vh * v
        llc = lookfrom - hor / 2.0 - ver / 2.0 - w
        return Camera(lookfrom.copy(), llc, hor, ver)

    def as12(self) -> np.ndarray:
        return np.concatenate([_f64(self.origin), _f64(self.lower_left_corner), _f64(self.horizontal),
                               _f64(self.vertical)])

    def c_struct(self) -> RtCamera:
        c = RtCamera()
        for name in ("origin", "lower_left_corner", "horizontal", "vertical"):
            v = _f64(getattr(self, name))
            getattr(c, name)[:] = [float(v[0]), float(v[1]), float(v[2])]
        return c


def make_params(width: int, height: int, spp: int, max_depth: int = 50, seed: int = 0, tmin: float = 0.0,
                jitter: bool = True, early_out: bool = True, scan_mode: int = SCAN_AUTO, shard_rank: int = 0,
                shard_count: int = 1, paths_per_lane: int = 0, chunks: int = 0, cull_smem: bool = False, variant: int = 0,
                albedo: float | None = None, sky_a=None, sky_b=None, scatter_mode: int | None = None) -> RtParams:
    """rt_params.  albedo / sky_a / sky_b / scatter_mode (any of them given) switch custom_shading on; left alone, the
    render uses the reference's constants (programs/main.cc:42,43,48)."""
    p = RtParams()
    _check(lib().rt_params_init(C.byref(p), width, height, spp, max_depth))
    if albedo is not None or sky_a is not None or sky_b is not None or scatter_mode is not None:
        p.custom_shading = 1
        if albedo is not None:
            p.albedo = albedo
        if sky_a is not None:
            p.sky_a[:] = [float(x) for x in sky_a]
        if sky_b is not None:
            p.sky_b[:] = [float(x) for x in sky_b]
        if scatter_mode is not None:
            p.scatter_mode = scatter_mode
    p.seed, p.tmin = seed & 0xFFFFFFFFFFFFFFFF, tmin
    p.jitter, p.early_out, p.scan_mode = int(jitter), int(early_out), scan_mode
    p.shard_rank, p.shard_count = shard_rank, shard_count
    p.reserved[0] = paths_per_lane  # tuning knob: paths per lane (0 = default)
    p.reserved[1] = chunks          # tuning knob: sample chunks per tile (0 = auto)
    # A/B knobs: 1 = cull array from TMA-staged shared memory instead of the constant bank; 2 = BVH mode through the
    # round-1 one-path-per-lane traversal kernel; 3 = BVH mode without the tie-grid fast path
    p.reserved[2] = variant if variant else (1 if cull_smem else 0)
    return p


class Scene:
    """Device-resident flatten of a hittable_list of spheres (list order preserved)."""

    def __init__(self, centres, radii, device: int = 0):
        self.centres = _f64(centres).reshape(-1, 3)
        self.radii = _f64(radii).reshape(-1)
        if len(self.centres) != len(self.radii):
            raise ValueError("centres and radii disagree")
        self.device = device
        self._h = C.c_void_p()
        _check(lib().rt_upload_scene(_dptr(self.centres), _dptr(self.radii), len(self.radii), device, C.byref(self._h)))

    @property
    def handle(self) -> C.c_void_p:
        if not self._h:
            raise RtError("scene is closed")
        return self._h

    def __len__(self) -> int:
        return len(self.radii)

    def update(self, centres, radii, refit: bool = True) -> None:
        """rt_update_scene: same spheres, moved / resized.  refit keeps the BVH topology (boxes recomputed)."""
        c, r = _f64(centres).reshape(-1, 3), _f64(radii).reshape(-1)
        if len(c) != len(self.radii) or len(r) != len(self.radii):
            raise ValueError("update keeps the sphere count")
        _check(lib().rt_update_scene(self.handle, _dptr(c), _dptr(r), len(r), int(refit)))
        self.centres, self.radii = c, r

    def close(self) -> None:
        if self._h:
            lib().rt_free_scene(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def render(scene: Scene, cam: Camera, params: RtParams, want_sums: bool = False, out: np.ndarray | None = None):
    """rt_render: host buffers in/out (the call a reference user makes instead of main.cc:72-88).  `out`: an
    (H, W, 4) uint8 C-contiguous array to receive the frame (e.g. a view of pinned memory), else a new array."""
    W, H = params.width, params.height
    if out is None:
        rgba = np.empty((H, W, 4), dtype=np.uint8)
    else:
        if out.shape != (H, W, 4) or out.dtype != np.uint8 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous (H, W, 4) uint8 array")
        rgba = out
    sums = np.empty((H, W, 3), dtype=np.float64) if want_sums else None
    st = RtStats()
    cs = cam.c_struct()
    _check(lib().rt_render(scene.handle, C.byref(cs), C.byref(params), rgba.ctypes.data_as(C.POINTER(C.c_uint8)),
                           _dptr(sums) if want_sums else None, C.byref(st)))
    return rgba, sums, st.as_dict()


def render_pass(scene: Scene, cam: Camera, params: RtParams, sample_begin: int, accum: np.ndarray | None = None,
                want_rgba: bool = True):
    """rt_render_pass: trace samples [sample_begin, sample_begin + params.spp) and add them to `accum` (H x W x 3
    uint64 fixed-point sums; None = start from zero).  Returns (accum, rgba over all samples so far, stats)."""
    W, H = params.width, params.height
    if accum is None:
        accum = np.zeros((H, W, 3), dtype=np.uint64)
    if accum.shape != (H, W, 3) or accum.dtype != np.uint64 or not accum.flags.c_contiguous:
        raise ValueError("accum must be a C-contiguous (H, W, 3) uint64 array")
    rgba = np.empty((H, W, 4), dtype=np.uint8) if want_rgba else None
    st = RtStats()
    cs = cam.c_struct()
    _check(lib().rt_render_pass(scene.handle, C.byref(cs), C.byref(params), sample_begin,
                                accum.ctypes.data_as(C.POINTER(C.c_uint64)),
                                rgba.ctypes.data_as(C.POINTER(C.c_uint8)) if want_rgba else None, C.byref(st)))
    return accum, rgba, st.as_dict()


class Accumulator:
    """rt_accum: device-resident progressive accumulator (fixed-point radiance sums + the frame of all samples so far)."""

    def __init__(self, width: int, height: int, device: int = 0):
        self.width, self.height, self.device = width, height, device
        self._h = C.c_void_p()
        _check(lib().rt_accum_create(width, height, device, C.byref(self._h)))

    @property
    def samples(self) -> int:
        return lib().rt_accum_samples(self._h)

    def add(self, scene: Scene, cam: Camera, params: RtParams, want_stats: bool = False):
        """Traces the next params.spp samples of every pixel; asynchronous unless want_stats."""
        st = RtStats()
        cs = cam.c_struct()
        _check(lib().rt_accum_add(scene.handle, C.byref(cs), C.byref(params), self._h, C.byref(st) if want_stats else None))
        return st.as_dict() if want_stats else None

    def frame(self) -> np.ndarray:
        rgba = np.empty((self.height, self.width, 4), dtype=np.uint8)
        _check(lib().rt_accum_frame(self._h, rgba.ctypes.data_as(C.POINTER(C.c_uint8))))
        return rgba

    def read(self) -> np.ndarray:
        sums = np.empty((self.height, self.width, 3), dtype=np.uint64)
        _check(lib().rt_accum_read(self._h, sums.ctypes.data_as(C.POINTER(C.c_uint64))))
        return sums

    def write(self, sums: np.ndarray, samples_done: int) -> None:
        sums = np.ascontiguousarray(sums, dtype=np.uint64)
        if sums.shape != (self.height, self.width, 3):
            raise ValueError("sums must be (H, W, 3) uint64")
        _check(lib().rt_accum_write(self._h, sums.ctypes.data_as(C.POINTER(C.c_uint64)), samples_done))

    def reset(self) -> None:
        _check(lib().rt_accum_reset(self._h))

    def close(self) -> None:
        if self._h:
            lib().rt_accum_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def render_multi(scenes: list, cam: Camera, params: RtParams, out: np.ndarray | None = None):
    """rt_render_multi: one frame over several device scenes (same spheres, one Scene per device) of this process."""
    W, H = params.width, params.height
    rgba = np.empty((H, W, 4), dtype=np.uint8) if out is None else out
    arr = (C.c_void_p * len(scenes))(*[s.handle for s in scenes])
    st = RtStats()
    cs = cam.c_struct()
    _check(lib().rt_render_multi(arr, len(scenes), C.byref(cs), C.byref(params), rgba.ctypes.data_as(C.POINTER(C.c_uint8)),
                                 C.byref(st)))
    return rgba, st.as_dict()


def accum_to_frame(params: RtParams, d_accum: int, total_samples: int, d_rgba: int, device: int, stream: int = 0) -> None:
    _check(lib().rt_accum_to_frame(C.byref(params), C.c_void_p(d_accum), total_samples, C.c_void_p(d_rgba), device,
                                   C.c_void_p(stream) if stream else None))


def render_pass_device(scene: Scene, cam: Camera, params: RtParams, sample_begin: int, d_accum: int, d_rgba: int = 0,
                       stream: int = 0) -> None:
    cs = cam.c_struct()
    _check(lib().rt_render_pass_device(scene.handle, C.byref(cs), C.byref(params), sample_begin, C.c_void_p(d_accum),
                                       C.c_void_p(d_rgba) if d_rgba else None, C.c_void_p(stream) if stream else None))


def render_device(scene: Scene, cam: Camera, params: RtParams, d_rgba: int, d_sums: int = 0, stream: int = 0) -> None:
    cs = cam.c_struct()
    _check(lib().rt_render_device(scene.handle, C.byref(cs), C.byref(params), C.c_void_p(d_rgba),
                                  C.c_void_p(d_sums) if d_sums else None, C.c_void_p(stream) if stream else None))


def render_finish(scene: Scene) -> dict:
    st = RtStats()
    _check(lib().rt_render_finish(scene.handle, C.byref(st)))
    return st.as_dict()


def tile_layout(params: RtParams) -> RtTileLayout:
    L = RtTileLayout()
    _check(lib().rt_get_tile_layout(C.byref(params), C.byref(L)))
    return L


def deinterleave(params: RtParams, d_gathered: int, d_rgba: int, device: int, stream: int = 0) -> None:
    _check(lib().rt_deinterleave(C.byref(params), C.c_void_p(d_gathered), C.c_void_p(d_rgba), device,
                                 C.c_void_p(stream) if stream else None))


def primary_hits(scene: Scene, cam: Camera, width: int, height: int, scan_mode: int = SCAN_AUTO, tmin: float = 0.0):
    p = make_params(width, height, 1, scan_mode=scan_mode, tmin=tmin)
    idx = np.empty((height, width), dtype=np.int32)
    t = np.empty((height, width), dtype=np.float64)
    cs = cam.c_struct()
    _check(lib().rt_primary_hits(scene.handle, C.byref(cs), C.byref(p), idx.ctypes.data_as(C.POINTER(C.c_int32)), _dptr(t)))
    return idx, t


def hit(scene: Scene, org, dirs, tmin: float = 0.0, tmax: float = float("inf"), scan_mode: int = SCAN_AUTO):
    org, dirs = _f64(org).reshape(-1, 3), _f64(dirs).reshape(-1, 3)
    n = len(org)
    idx = np.empty(n, dtype=np.int32)
    rec = np.empty((n, 8), dtype=np.float64)
    _check(lib().rt_hit(scene.handle, _dptr(org), _dptr(dirs), n, tmin, tmax, scan_mode,
                        idx.ctypes.data_as(C.POINTER(C.c_int32)), _dptr(rec)))
    return idx, rec


def ray_color(scene: Scene, org, dirs, depth: int, seed: int = 0, early_out: bool = False, scan_mode: int = SCAN_AUTO):
    org, dirs = _f64(org).reshape(-1, 3), _f64(dirs).reshape(-1, 3)
    n = len(org)
    rgb = np.empty((n, 3), dtype=np.float64)
    st = RtStats()
    _check(lib().rt_ray_color(scene.handle, _dptr(org), _dptr(dirs), n, depth, seed & 0xFFFFFFFFFFFFFFFF, int(early_out),
                              scan_mode, _dptr(rgb), C.byref(st)))
    return rgb, st.as_dict()


def ray_color_params(scene: Scene, org, dirs, params: RtParams):
    """rt_ray_color_params: ray_color with tmin / shading / depth / seed / early_out / scan_mode from `params`."""
    org, dirs = _f64(org).reshape(-1, 3), _f64(dirs).reshape(-1, 3)
    n = len(org)
    rgb = np.empty((n, 3), dtype=np.float64)
    st = RtStats()
    _check(lib().rt_ray_color_params(scene.handle, _dptr(org), _dptr(dirs), n, C.byref(params), _dptr(rgb), C.byref(st)))
    return rgb, st.as_dict()


def write_color(rgb_sum, spp: int, device: int = 0) -> np.ndarray:
    rgb_sum = _f64(rgb_sum).reshape(-1, 3)
    out = np.empty((len(rgb_sum), 3), dtype=np.int32)
    _check(lib().rt_write_color(_dptr(rgb_sum), len(rgb_sum), spp, device, out.ctypes.data_as(C.POINTER(C.c_int32))))
    return out


def get_ray(cam: Camera, uv, device: int = 0) -> np.ndarray:
    uv = _f64(uv).reshape(-1, 2)
    out = np.empty((len(uv), 6), dtype=np.float64)
    cs = cam.c_struct()
    _check(lib().rt_get_ray(C.byref(cs), _dptr(uv), len(uv), device, _dptr(out)))
    return out


def philox(ctr, key, device: int = 0) -> np.ndarray:
    ctr = np.ascontiguousarray(ctr, dtype=np.uint32).reshape(-1, 4)
    key = np.ascontiguousarray(key, dtype=np.uint32).reshape(2)
    out = np.empty_like(ctr)
    u32p = C.POINTER(C.c_uint32)
    _check(lib().rt_philox(ctr.ctypes.data_as(u32p), key.ctypes.data_as(u32p), len(ctr), device, out.ctypes.data_as(u32p)))
    return out


def measure_fp32_peak(device: int = 0) -> tuple[float, float]:
    """(FMA lanes per second, ms of the timed launch) from the FFMA microbenchmark kernel."""
    f, ms = C.c_double(), C.c_double()
    _check(lib().rt_measure_fp32_peak(device, C.byref(f), C.byref(ms)))
    return f.value, ms.value


def check_division(n: int, seed: int = 1, device: int = 0) -> int:
    """Mismatches between the short hit-distance division and __ddiv_rn on n random operand pairs (must be 0)."""
    bad = C.c_uint64()
    _check(lib().rt_check_division(device, n, seed, C.byref(bad)))
    return bad.value


def device_info(device: int = 0) -> dict:
    sm, maj, mnr = C.c_int32(), C.c_int32(), C.c_int32()
    name = C.create_string_buffer(256)
    _check(lib().rt_device_info(device, C.byref(sm), C.byref(maj), C.byref(mnr), name, 256))
    return {"sm_count": sm.value, "cc": (maj.value, mnr.value), "name": name.value.decode()}
