"""The C++ host-side mirror (grove_b200/csrc/host): producer vectors transcribed from the reference's
syncflow_test.go, PodGang -> table encoding, and (gpu) GpuBackend cycles through the C ABI."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_test_bin(built_lib, tmp_path_factory):
    out = os.path.join(ROOT, "tests", "cpp", "test_host.bin")
    srcs = [os.path.join(ROOT, "tests", "cpp", "test_host.cpp"), os.path.join(ROOT, "grove_b200", "csrc", "host", "grove_host.cpp")]
    libdir = os.path.join(ROOT, "grove_b200")
    env = dict(os.environ); env.pop("CC", None); env.pop("CXX", None)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-pthread", "-o", out, *srcs, f"-L{libdir}", "-lgrove_place",
                           f"-Wl,-rpath,{libdir}", "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"], env=env)
    return out


def test_host_mirror_cpu(host_test_bin):
    r = subprocess.run([host_test_bin, "cpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "host tests ok" in r.stdout


@pytest.mark.gpu
def test_host_mirror_gpu(host_test_bin):
    r = subprocess.run([host_test_bin, "gpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "host tests ok (cpu+gpu)" in r.stdout
