"""GPU: the C++ drop-in classes (atlas_engine_b200/host/AtlasRT.h) driven like the engine drives the originals —
concurrent BLAS builds from several threads, TLAS, MeshData/RayTracingWorld mirrors, trace — against oracle/_ref."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_cpp_parity(ctx, tmp_path):
    ref = os.path.join(ROOT, "oracle", "_ref", "libatlas_ref.so")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/libatlas_ref.so not available")
    exe = str(tmp_path / "host_parity")
    lib_dir = os.path.join(ROOT, "atlas_engine_b200")
    subprocess.run(["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "cpp", "host_parity.cpp"), "-o", exe,
                    "-L" + lib_dir, "-latlas_rt", "-Wl,-rpath," + lib_dir, "-ldl", "-lpthread"], check=True)
    r = subprocess.run([exe, ref], capture_output=True, text=True, timeout=200, cwd=ROOT)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and "ALL OK" in r.stdout
