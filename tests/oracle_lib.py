"""ctypes access to the test-only oracle libraries (oracle/liboracle.so, oracle/_ref/libref.so).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this.
Both libraries expose the same entry points (prefix orc_ / ref_), so most helpers take `which`.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(REPO, "oracle", "liboracle.so")
REF_SO = os.path.join(REPO, "oracle", "_ref", "libref.so")

RNG_RAND15, RNG_PHILOX = 0, 1

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)
u8p = C.POINTER(C.c_uint8)
u64p = C.POINTER(C.c_uint64)


class OrcStats(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("samples", "casts", "black", "primary_hits", "early_outs")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class OrcShading(C.Structure):
    """orc_shading / ref_shading: the constants ray_color hard-codes (programs/main.cc:40,42,43,48)."""
    _fields_ = [("tmin", C.c_double), ("albedo", C.c_double), ("sky_a", C.c_double * 3), ("sky_b", C.c_double * 3),
                ("scatter_mode", C.c_int)]


SCATTER_HEMISPHERE, SCATTER_LAMBERTIAN = 0, 1


def shading(tmin=0.0, albedo=0.5, sky_a=(1.0, 1.0, 1.0), sky_b=(0.5, 0.7, 1.0), scatter_mode=SCATTER_HEMISPHERE) -> OrcShading:
    sh = OrcShading()
    sh.tmin, sh.albedo, sh.scatter_mode = tmin, albedo, scatter_mode
    sh.sky_a[:] = list(sky_a)
    sh.sky_b[:] = list(sky_b)
    return sh


_cache = {}


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def oracle() -> C.CDLL:
    if "orc" not in _cache:
        if not os.path.exists(ORACLE_SO):
            import subprocess
            subprocess.run(["make", "-s", "-C", os.path.join(REPO, "oracle"), "oracle"], check=True)
        L = C.CDLL(ORACLE_SO)
        L.orc_main_ppm.restype = C.c_long
        L.orc_main_ppm.argtypes = [C.c_uint64, C.c_char_p, C.c_long]
        L.orc_render_rows.argtypes = [dp, dp, C.c_int, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_int, u8p, dp, C.POINTER(OrcStats)]
        L.orc_render_rows_ex.argtypes = [dp, dp, C.c_int, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int,
                                         C.POINTER(OrcShading), C.c_int, C.c_int, C.c_int, u8p, dp, C.POINTER(OrcStats)]
        L.orc_ray_color_batch_ex.argtypes = [dp, dp, C.c_int, dp, dp, u64p, C.c_int, C.c_int, C.POINTER(OrcShading), C.c_int,
                                             C.c_int, dp, C.POINTER(OrcStats)]
        L.orc_primary_hits.argtypes = [dp, dp, C.c_int, dp, C.c_int, C.c_int, ip, dp]
        L.orc_hit_batch.argtypes = [dp, dp, C.c_int, dp, dp, C.c_int, C.c_double, C.c_double, ip, dp]
        L.orc_sphere_hit_batch.argtypes = [dp, dp, dp, dp, C.c_int, C.c_double, C.c_double, ip, dp]
        L.orc_ray_color_batch.argtypes = [dp, dp, C.c_int, dp, dp, u64p, C.c_int, C.c_int, C.c_int, C.c_int, dp,
                                          C.POINTER(OrcStats)]
        L.orc_get_ray_batch.argtypes = [dp, dp, C.c_int, dp]
        L.orc_default_camera.argtypes = [dp, dp]
        L.orc_write_color_batch.argtypes = [dp, C.c_int, C.c_int, ip]
        L.orc_random_in_hemisphere_batch.argtypes = [dp, u64p, C.c_int, dp]
        L.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.orc_max_threads.restype = C.c_int
        _cache["orc"] = L
    return _cache["orc"]


def ref(opt: str = "") -> C.CDLL:
    """oracle/_ref/libref.so; opt="O0": the same sources built without optimisation (the reference's own CMake
    configuration, programs/CMakeLists.txt:1-6), used only for bench.py's cpu_baseline of config 1."""
    key = "ref" + opt
    if key not in _cache:
        L = C.CDLL(REF_SO if not opt else REF_SO.replace("libref.so", f"libref_{opt}.so"))
        L.ref_main_ppm.restype = C.c_long
        L.ref_main_ppm.argtypes = [C.c_uint64, C.c_char_p, C.c_long]
        L.ref_render_rows.argtypes = [dp, dp, C.c_int, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int,
                                      C.c_int, u8p, dp]
        L.ref_primary_hits.argtypes = [dp, dp, C.c_int, dp, C.c_int, C.c_int, ip, dp]
        L.ref_hit_batch.argtypes = [dp, dp, C.c_int, dp, dp, C.c_int, C.c_double, C.c_double, ip, dp]
        L.ref_sphere_hit_batch.argtypes = [dp, dp, dp, dp, C.c_int, C.c_double, C.c_double, ip, dp]
        L.ref_ray_color_batch.argtypes = [dp, dp, C.c_int, dp, dp, u64p, C.c_int, C.c_int, dp]
        L.ref_ray_color_param_batch.argtypes = [dp, dp, C.c_int, dp, dp, u64p, C.POINTER(OrcShading), C.c_int, C.c_int, dp]
        L.ref_render_rows_param.argtypes = [dp, dp, C.c_int, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64,
                                            C.POINTER(OrcShading), C.c_int, C.c_int, C.c_int, u8p, dp]
        L.ref_get_ray_batch.argtypes = [dp, dp, C.c_int, dp]
        L.ref_default_camera.argtypes = [dp, dp]
        L.ref_write_color_batch.argtypes = [dp, C.c_int, C.c_int, ip]
        L.ref_random_in_hemisphere_batch.argtypes = [dp, u64p, C.c_int, dp]
        L.ref_max_threads.restype = C.c_int
        _cache[key] = L
    return _cache[key]


def have_ref_O0() -> bool:
    return os.path.exists(REF_SO.replace("libref.so", "libref_O0.so"))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(dp)


def main_ppm(which: str, seed: int = 0x9E3779B97F4A7C15) -> bytes:
    buf = C.create_string_buffer(4 << 20)
    fn = oracle().orc_main_ppm if which == "orc" else ref().ref_main_ppm
    n = fn(seed, buf, len(buf))
    assert n > 0
    return buf.raw[:n]


def render(which: str, centres, radii, cam12, W, H, spp, max_depth=50, seed=0, rng_mode=RNG_RAND15, early_out=False,
           j0=0, j1=None, nthreads=0, want_sums=False, shading: OrcShading | None = None):
    """Returns (rgb (H,W,3) uint8, sums or None, stats dict).  shading=None: the reference's constants."""
    centres, radii, cam12 = _f64(centres).reshape(-1, 3), _f64(radii), _f64(cam12)
    j1 = H if j1 is None else j1
    rgb = np.zeros((H, W, 3), dtype=np.uint8)
    if which == "orc":
        sums = np.zeros((H, W, 3), dtype=np.float64) if want_sums else None
        st = OrcStats()
        oracle().orc_render_rows_ex(_p(centres), _p(radii), len(radii), _p(cam12), W, H, spp, max_depth, seed, rng_mode,
                                    int(early_out), C.byref(shading) if shading is not None else None, j0, j1, nthreads,
                                    rgb.ctypes.data_as(u8p), _p(sums) if want_sums else None, C.byref(st))
        return rgb, sums, st.as_dict()
    assert rng_mode == RNG_RAND15 and not early_out and not want_sums
    stats = np.zeros(3, dtype=np.float64)
    if which == "refO0":
        assert shading is None
        ref("O0").ref_render_rows(_p(centres), _p(radii), len(radii), _p(cam12), W, H, spp, max_depth, seed, j0, j1, nthreads,
                                  rgb.ctypes.data_as(u8p), _p(stats))
        return rgb, None, {"samples": stats[0], "casts": stats[1], "black": stats[2]}
    if shading is not None:
        ref().ref_render_rows_param(_p(centres), _p(radii), len(radii), _p(cam12), W, H, spp, max_depth, seed,
                                    C.byref(shading), j0, j1, nthreads, rgb.ctypes.data_as(u8p), _p(stats))
        return rgb, None, {"samples": stats[0], "casts": stats[1], "black": stats[2]}
    ref().ref_render_rows(_p(centres), _p(radii), len(radii), _p(cam12), W, H, spp, max_depth, seed, j0, j1, nthreads,
                          rgb.ctypes.data_as(u8p), _p(stats))
    return rgb, None, {"samples": stats[0], "casts": stats[1], "black": stats[2]}


def primary_hits(which, centres, radii, cam12, W, H):
    centres, radii, cam12 = _f64(centres).reshape(-1, 3), _f64(radii), _f64(cam12)
    idx = np.empty((H, W), dtype=np.int32)
    t = np.empty((H, W), dtype=np.float64)
    fn = oracle().orc_primary_hits if which == "orc" else ref().ref_primary_hits
    fn(_p(centres), _p(radii), len(radii), _p(cam12), W, H, idx.ctypes.data_as(ip), _p(t))
    return idx, t


def hit_batch(which, centres, radii, org, dirs, tmin=0.0, tmax=float("inf")):
    centres, radii = _f64(centres).reshape(-1, 3), _f64(radii)
    org, dirs = _f64(org).reshape(-1, 3), _f64(dirs).reshape(-1, 3)
    n = len(org)
    idx = np.empty(n, dtype=np.int32)
    rec = np.empty((n, 8), dtype=np.float64)
    fn = oracle().orc_hit_batch if which == "orc" else ref().ref_hit_batch
    fn(_p(centres), _p(radii), len(radii), _p(org), _p(dirs), n, tmin, tmax, idx.ctypes.data_as(ip), _p(rec))
    return idx, rec


def sphere_hit_batch(which, centre, radius, org, dirs, tmin=0.0, tmax=float("inf")):
    centre, radius = _f64(centre).reshape(-1, 3), _f64(radius)
    org, dirs = _f64(org).reshape(-1, 3), _f64(dirs).reshape(-1, 3)
    n = len(org)
    hit = np.empty(n, dtype=np.int32)
    rec = np.empty((n, 8), dtype=np.float64)
    fn = oracle().orc_sphere_hit_batch if which == "orc" else ref().ref_sphere_hit_batch
    fn(_p(centre), _p(radius), _p(org), _p(dirs), n, tmin, tmax, hit.ctypes.data_as(ip), _p(rec))
    return hit, rec


def ray_color_batch(which, centres, radii, org, dirs, seeds, depth, rng_mode=RNG_RAND15, early_out=False,
                    shading: OrcShading | None = None):
    centres, radii = _f64(centres).reshape(-1, 3), _f64(radii)
    org, dirs = _f64(org).reshape(-1, 3), _f64(dirs).reshape(-1, 3)
    seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
    n = len(org)
    rgb = np.empty((n, 3), dtype=np.float64)
    if which == "orc":
        st = OrcStats()
        oracle().orc_ray_color_batch_ex(_p(centres), _p(radii), len(radii), _p(org), _p(dirs), seeds.ctypes.data_as(u64p),
                                        rng_mode, int(early_out), C.byref(shading) if shading is not None else None, n,
                                        depth, _p(rgb), C.byref(st))
        return rgb, st.as_dict()
    assert rng_mode == RNG_RAND15 and not early_out
    if shading is not None:
        ref().ref_ray_color_param_batch(_p(centres), _p(radii), len(radii), _p(org), _p(dirs), seeds.ctypes.data_as(u64p),
                                        C.byref(shading), n, depth, _p(rgb))
        return rgb, None
    ref().ref_ray_color_batch(_p(centres), _p(radii), len(radii), _p(org), _p(dirs), seeds.ctypes.data_as(u64p), n, depth,
                              _p(rgb))
    return rgb, None


def get_ray_batch(which, cam12, uv):
    cam12, uv = _f64(cam12), _f64(uv).reshape(-1, 2)
    out = np.empty((len(uv), 6), dtype=np.float64)
    fn = oracle().orc_get_ray_batch if which == "orc" else ref().ref_get_ray_batch
    fn(_p(cam12), _p(uv), len(uv), _p(out))
    return out


def default_camera(which):
    cam12 = np.empty(12, dtype=np.float64)
    aspect = C.c_double()
    fn = oracle().orc_default_camera if which == "orc" else ref().ref_default_camera
    fn(_p(cam12), C.cast(C.byref(aspect), dp))
    return cam12, aspect.value


def write_color_batch(which, rgb_sum, spp):
    rgb_sum = _f64(rgb_sum).reshape(-1, 3)
    out = np.empty((len(rgb_sum), 3), dtype=np.int32)
    fn = oracle().orc_write_color_batch if which == "orc" else ref().ref_write_color_batch
    fn(_p(rgb_sum), len(rgb_sum), spp, out.ctypes.data_as(ip))
    return out


def random_in_hemisphere_batch(which, normals, seeds):
    normals = _f64(normals).reshape(-1, 3)
    seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
    out = np.empty_like(normals)
    fn = oracle().orc_random_in_hemisphere_batch if which == "orc" else ref().ref_random_in_hemisphere_batch
    fn(_p(normals), seeds.ctypes.data_as(u64p), len(normals), _p(out))
    return out


def philox(ctr4, key2):
    out = (C.c_uint32 * 4)()
    oracle().orc_philox4x32_10((C.c_uint32 * 4)(*ctr4), (C.c_uint32 * 2)(*key2), out)
    return list(out)


def psnr(a: np.ndarray, b: np.ndarray) -> float:
    """PSNR of two 8-bit images (any matching shape)."""
    d = a.astype(np.float64) - b.astype(np.float64)
    mse = float(np.mean(d * d))
    return float("inf") if mse == 0 else 10.0 * np.log10(255.0 * 255.0 / mse)
