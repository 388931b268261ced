// The C++ mirror of RegisterRGBD360::RegisterDensePhotoICP (include/RegisterRGBD360_b200.hpp) through the C ABI:
//   rig_demo <dir> <rows> <cols>   with <dir>/rgb.bin (frame1 then frame2: 8 x rows x cols x 3 u8 each), depth.bin (u16), rt.bin (8 x 16 f32)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "RegisterRGBD360_b200.hpp"

template <typename T> static std::vector<T> slurp(const std::string& p, size_t n) {
    std::vector<T> v(n);
    FILE* f = std::fopen(p.c_str(), "rb");
    if (!f || std::fread(v.data(), sizeof(T), n, f) != n) { std::fprintf(stderr, "cannot read %s\n", p.c_str()); std::exit(2); }
    std::fclose(f);
    return v;
}

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const std::string dir = argv[1];
    const int rows = std::atoi(argv[2]), cols = std::atoi(argv[3]);
    const size_t npx = (size_t)8 * rows * cols;
    const std::vector<uint8_t> rgb = slurp<uint8_t>(dir + "/rgb.bin", 2 * npx * 3);
    const std::vector<uint16_t> depth = slurp<uint16_t>(dir + "/depth.bin", 2 * npx);
    const std::vector<float> Rt = slurp<float>(dir + "/rt.bin", 8 * 16);
    r360::RegisterRGBD360 reg(rows, cols, 0, 3);
    reg.setExtrinsics(Rt.data());
    const r360::RigFrame f1 = { rgb.data(), depth.data() }, f2 = { rgb.data() + npx * 3, depth.data() + npx };
    // as upstream: the guess comes back, with the rig's information matrix
    if (!reg.RegisterDensePhotoICP(f1, f2)) { std::fprintf(stderr, "faithful run returned false\n"); return 1; }
    for (int k = 0; k < 16; ++k)
        if (reg.getPose()[k] != ((k % 5 == 0) ? 1.f : 0.f)) { std::fprintf(stderr, "faithful run moved the pose\n"); return 1; }
    if (!(reg.getInfoMat()[0] > 0.f)) { std::fprintf(stderr, "no information matrix\n"); return 1; }
    // the intended loop: moves towards the other frame
    reg.setFaithfulNewError(false);
    if (!reg.RegisterDensePhotoICP(f1, f2)) { std::fprintf(stderr, "fixed run returned false\n"); return 1; }
    double t = 0;
    for (int k = 12; k < 15; ++k) t += (double)reg.getPose()[k] * reg.getPose()[k];
    if (!(std::sqrt(t) > 1e-3)) { std::fprintf(stderr, "fixed run did not move\n"); return 1; }
    std::printf("OK |t| = %.4f m\n", std::sqrt(t));
    return 0;
}
