"""include/rt_host.hpp (the C++ host API with the reference's names): host-only behaviour, no GPU needed."""
import os
import subprocess

from conftest import REPO


def test_flatten_and_api_surface(tmp_path, rt):
    exe = tmp_path / "flatten_test"
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    pkg = os.path.dirname(rt.LIB_PATH)
    subprocess.run([cxx, "-O1", "-std=c++17", "-pthread", "-I", os.path.join(REPO, "include"), "-I", os.path.join(REPO, "include", "compat"),
                    os.path.join(REPO, "tests", "cpp", "flatten_test.cc"), "-o", str(exe), "-L", pkg, "-lrt_b200",
                    f"-Wl,-rpath,{pkg}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stdout + out.stderr

    # rt::book_scene / rt::book_camera (C++) generate the doubles scenes.py generates
    import numpy as np
    from petershirleyraytracer_b200 import scenes
    lines = subprocess.run([str(exe), "book"], capture_output=True, text=True, check=True).stdout.split("\n")
    n = int(lines[0])
    rows = np.array([[float.fromhex(v) for v in ln.split()] for ln in lines[1:1 + n]])
    c, r = scenes.book_scene(11, 42)
    assert n == len(r) == 485
    assert np.array_equal(rows[:, :3], c) and np.array_equal(rows[:, 3], r)
    cam = np.array([[float.fromhex(v) for v in ln.split()] for ln in lines[1 + n:5 + n]]).reshape(-1)
    assert np.allclose(cam, scenes.book_camera(1200, 800).as12(), rtol=1e-14, atol=1e-15)


def test_bvh_builder_invariants(tmp_path):
    """rt_bvh.h on the host: each sphere in exactly one leaf, every (padded FP32) box holds the spheres below it -- the
    property the exact traversal rests on -- for SAH builds of ordinary and adversarial scenes and after a refit."""
    exe = tmp_path / "bvh_host_test"
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O1", "-std=c++17", "-pthread", "-I", os.path.join(REPO, "petershirleyraytracer_b200", "csrc"),
                    os.path.join(REPO, "tests", "cpp", "bvh_host_test.cc"), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout[-2000:] + out.stderr[-2000:]


def test_tie_grid_containment(tmp_path):
    """rt_bvh.h build_tie_grid on the host: every sphere that can pass the device's FP32 shell test at a point is a
    giant or listed in the cell the device looks up (ordinary, clustered, nested, coincident and tiny scenes)."""
    exe = tmp_path / "tie_grid_test"
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O1", "-std=c++17", "-pthread", "-I", os.path.join(REPO, "petershirleyraytracer_b200", "csrc"),
                    os.path.join(REPO, "tests", "cpp", "tie_grid_test.cc"), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout[-2000:] + out.stderr[-2000:]


def test_work_unit_plans_cover_every_sample_once(tmp_path):
    """rt_units.h on the host: whatever the frame, spp, grid, kernel and tuning knob, the graded chunks of a tile are
    contiguous, non-empty and add up to spp, and the unit ids of a launch enumerate every (tile, chunk) exactly once."""
    exe = tmp_path / "units_test"
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O1", "-std=c++17", "-I", os.path.join(REPO, "petershirleyraytracer_b200", "csrc"),
                    os.path.join(REPO, "tests", "cpp", "units_test.cc"), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout[-2000:] + out.stderr[-2000:]
