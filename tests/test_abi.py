"""The C-ABI library loads and exports every symbol include/rt.h declares; host-only entry points work
without a GPU; compute entry points fail loudly (RT_ERR_CUDA) instead of falling back to the CPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import REPO


def _header_functions():
    src = open(os.path.join(REPO, "include", "rt.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree(rt):
    assert _header_functions() == sorted(rt.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol(rt):
    L = rt.lib()
    for name in _header_functions():
        assert hasattr(L, name), name
    assert L.rt_abi_version() == rt.ABI_VERSION == 3


def test_struct_layouts(rt):
    assert C.sizeof(rt.RtCamera) == 96
    assert C.sizeof(rt.RtParams) == 128   # ABI 3: + custom_shading, scatter_mode, albedo, sky_a[3], sky_b[3]
    assert C.sizeof(rt.RtStats) == 96
    assert C.sizeof(rt.RtTileLayout) == 32


def test_params_init_fills_the_reference_constants(rt):
    """rt_params_init: tmin 0 (main.cc:40), hemisphere scatter (:42), albedo 0.5 (:43), sky (:48); custom_shading off."""
    p = rt.make_params(400, 225, 100, 50)
    assert (p.width, p.height, p.spp, p.max_depth, p.tmin, p.jitter, p.shard_count) == (400, 225, 100, 50, 0.0, 1, 1)
    assert p.custom_shading == 0 and p.scatter_mode == rt.SCATTER_HEMISPHERE and p.albedo == 0.5
    assert list(p.sky_a) == [1.0, 1.0, 1.0] and list(p.sky_b) == [0.5, 0.7, 1.0]
    q = rt.make_params(400, 225, 100, 50, albedo=0.7, scatter_mode=rt.SCATTER_LAMBERTIAN)
    assert q.custom_shading == 1 and q.albedo == 0.7 and list(q.sky_b) == [0.5, 0.7, 1.0]
    for bad in (dict(albedo=1.5), dict(albedo=float("nan")), dict(sky_a=(2.0, 0, 0)), dict(scatter_mode=7)):
        with pytest.raises(rt.RtError):
            rt.tile_layout(rt.make_params(64, 64, 1, **bad))


def test_tile_layout_host_only(rt):
    p = rt.make_params(1200, 800, 500, shard_count=8, shard_rank=3)
    L = rt.tile_layout(p)
    assert (L.tile_w, L.tile_h, L.tiles_x, L.tiles_y, L.tiles_total) == (8, 8, 150, 100, 15000)
    assert L.tiles_per_shard == 1875 and L.shard_bytes == 1875 * 64 * 4
    p = rt.make_params(401, 227, 1, shard_count=3, shard_rank=0)
    L = rt.tile_layout(p)
    assert (L.tiles_x, L.tiles_y, L.tiles_total, L.tiles_per_shard) == (51, 29, 1479, 493)


def test_bad_params_rejected(rt):
    with pytest.raises(rt.RtError):
        rt.tile_layout(rt.make_params(1, 10, 1))
    with pytest.raises(rt.RtError):
        rt.tile_layout(rt.make_params(10, 10, 0))
    with pytest.raises(rt.RtError):
        rt.tile_layout(rt.make_params(10, 10, 1, shard_count=2, shard_rank=2))


def test_no_cpu_fallback(rt):
    """Without a CUDA device the product path must raise, not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(rt.RtError) as e:
        rt.Scene(np.zeros((1, 3)), np.ones(1))
    assert "rt error -2" in str(e.value)


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may touch oracle/."""
    pkg = os.path.join(REPO, "petershirleyraytracer_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h", ".hpp")):
                text = open(os.path.join(root, f), errors="replace").read()
                assert "liboracle" not in text and "libref" not in text and "oracle_lib" not in text, f
    for root, _, files in os.walk(os.path.join(REPO, "include")):
        for f in files:
            text = open(os.path.join(root, f)).read()
            assert "liboracle" not in text and "rt_oracle" not in text, f
