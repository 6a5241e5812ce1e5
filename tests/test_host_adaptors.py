"""The C++ host side (host/slamb200_adaptors.hpp: myslam::ORBextractor & co. over the C ABI) compiles without
OpenCV, links against libslamb200.so, and fails loudly without a GPU; with a GPU it runs the reference's calls."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKGD = os.path.join(ROOT, "a-simple-stereo-slam-system-with-deep-loop-closing_b200")


def build(tmp):
    exe = os.path.join(tmp, "host_adaptor_test")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "host_adaptor_test.cpp"),
                           "-L" + PKGD, "-lslamb200", "-Wl,-rpath," + PKGD])
    return exe


def test_adaptors_compile_and_report_missing_device(tmp_path, pkg):
    exe = build(str(tmp_path))
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr


@pytest.mark.gpu
def test_adaptors_run_the_reference_call_sequence(tmp_path, pkg):
    exe = build(str(tmp_path))
    out = subprocess.run([exe, "gpu"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "DetectAndCompute" in out.stdout
