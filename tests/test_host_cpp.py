"""include/rt_host.hpp (the C++ host API with the reference's names): host-only behaviour, no GPU needed."""
import os
import subprocess

from conftest import REPO


def test_flatten_and_api_surface(tmp_path, rt):
    exe = tmp_path / "flatten_test"
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    pkg = os.path.dirname(rt.LIB_PATH)
    subprocess.run([cxx, "-O1", "-std=c++17", "-I", os.path.join(REPO, "include"), "-I", os.path.join(REPO, "include", "compat"),
                    os.path.join(REPO, "tests", "cpp", "flatten_test.cc"), "-o", str(exe), "-L", pkg, "-lrt_b200",
                    f"-Wl,-rpath,{pkg}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stdout + out.stderr
