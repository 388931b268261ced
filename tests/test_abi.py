"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/r360.h declares, fails loudly without a GPU, and the C++ mirror header compiles."""
import ctypes as C
import os
import re
import subprocess
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def native(r360):
    if not os.path.exists(os.path.join(ROOT, "rgbd360_b200", "librgbd360_b200.so")):
        r360.build_native()
    return r360


def declared_symbols():
    h = open(os.path.join(ROOT, "include", "r360.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(r360_[a-z0-9_]+)\s*\(", h)))


def test_library_exports_every_declared_symbol(native):
    L = native.lib()
    syms = declared_symbols()
    assert len(syms) >= 20
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(native.native.EXPORTS) == syms


def test_struct_layouts_match_header(native):
    assert C.sizeof(native.Params) == 64
    assert C.sizeof(native.Result) == 336
    assert C.sizeof(native.IterRecord) == 224
    assert native.native.RESULT_DTYPE.itemsize == 336
    p = native.default_params()
    assert (p.n_levels, p.max_iters, p.method, p.occlusion, p.n_sensors_mask) == (4, 10, 2, 0, 8)
    assert p.std_photo == np.float32(6.0 / 255) and p.min_depth == np.float32(0.3) and p.tol_residual == 1e-3


def test_no_cpu_fallback(native):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(native.R360Error, match="no CUDA device|no CPU fallback"):
        native.Context(64, 128, 2, 1)


def test_product_does_not_reference_oracle():
    """Nothing under rgbd360_b200/ or include/ may import, link or call oracle/."""
    bad = []
    for base in ("rgbd360_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            for f in fs:
                if f.endswith((".py", ".h", ".hpp", ".cu", ".cuh", ".cpp", "Makefile")):
                    t = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"(from|import)\s+oracle|oracle/|librpi_oracle|orc_[a-z]+\(", t):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_cpp_mirror_header_compiles(native, tmp_path):
    """include/RegisterPhotoICP_b200.hpp (the reference's class surface) compiles and links."""
    src = tmp_path / "t.cpp"
    src.write_text('''
#include "RegisterPhotoICP_b200.hpp"
int main() {
    RegisterPhotoICP reg;
    reg.setNumPyr(3); reg.setMinDepth(0.3f); reg.setMaxDepth(6.f); reg.setGrayVariance(3.f/255); reg.setDepthVariance(0.2f);
    reg.setVisualization(false); reg.useSaliency(false);
    return reg.nPyrLevels == 3 ? 0 : 1;
}
''')
    exe = tmp_path / "t"
    subprocess.check_call(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", os.path.join(ROOT, "rgbd360_b200"), "-lrgbd360_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "rgbd360_b200")])
    assert subprocess.call([str(exe)]) == 0


def test_cpp_rig_and_driver_headers_compile(native, tmp_path):
    """include/RegisterRGBD360_b200.hpp and include/Drivers360_b200.hpp compile and link against the C ABI; without a GPU
    the rig class fails loudly at construction (r360_create: no CUDA device), it never falls back."""
    src = tmp_path / "t.cpp"
    src.write_text('''
#include <cstdio>
#include <stdexcept>
#include "RegisterRGBD360_b200.hpp"
#include "Drivers360_b200.hpp"
int main() {
    const r360::Pose I = r360::identityPose();
    if (r360::toRobotFrame(r360::toSphereFrame(I))[0] < 0.999f) return 2;
    try {
        r360::RegisterRGBD360 reg(240, 320);
        reg.setFaithfulNewError(false);
    } catch (const std::runtime_error& e) { std::printf("no device: %s\\n", e.what()); return 0; }
    return 0;
}
''')
    exe = tmp_path / "t"
    subprocess.check_call(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", os.path.join(ROOT, "rgbd360_b200"), "-lrgbd360_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "rgbd360_b200")])
    assert subprocess.call([str(exe)]) == 0
