"""The multi-GPU exchange of the path from a C++ host: r360_allgather_results over NCCL, one host
thread + ctx + rank per GPU (tests/cpp/allgather_demo.cpp).  The CPU part checks that the demo builds against the
C ABI and the system NCCL and that the entry point validates its arguments; the run itself needs two GPUs."""
import os
import subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def build_demo(tmp_path):
    exe = tmp_path / "allgather_demo"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"),
                           os.path.join(ROOT, "tests", "cpp", "allgather_demo.cpp"), "-o", str(exe),
                           "-L", os.path.join(ROOT, "rgbd360_b200"), "-lrgbd360_b200", "-lnccl",
                           "-L", os.path.join(CUDA, "lib64"), "-lcudart", "-lpthread",
                           "-Wl,-rpath," + os.path.join(ROOT, "rgbd360_b200")])
    return exe


def test_allgather_demo_builds_and_entry_point_checks_arguments(r360, tmp_path):
    if not os.path.exists("/usr/include/nccl.h"):
        pytest.skip("no system NCCL headers")
    build_demo(tmp_path)
    L = r360.lib()
    assert L.r360_allgather_results(None, None, None, 0, 1, None) == -1          # R360_E_ARG: no ctx


@pytest.mark.gpu
def test_cpp_allgather_two_gpus(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    exe = build_demo(tmp_path)
    out = subprocess.run([str(exe), "2", "4"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "allgather_demo ok" in out.stdout, out.stdout + out.stderr
