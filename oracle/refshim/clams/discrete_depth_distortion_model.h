// Stand-in for CLAMS' depth distortion model (Calib360.h:39, 46, 98-111): Calib360 holds one per sensor and loads
// them in loadIntrinsicCalibration(), which the stitch harness never calls.  *** TEST INFRASTRUCTURE ONLY. ***
#pragma once
#include <string>
#include <iostream>
#include <cstdlib>
namespace clams {
class DiscreteDepthDistortionModel {
public:
    static void na() { std::cerr << "refshim: clams::DiscreteDepthDistortionModel is not implemented\n"; abort(); }
    void load(const std::string&) { na(); }
    void downsampleParams(int) { na(); }
    template <typename M> void undistort(M*) const { na(); }
    std::string status() const { return "refshim stand-in"; }
};
}  // namespace clams
