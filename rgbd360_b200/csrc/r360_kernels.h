// r360_kernels.h -- launch interface between the host API (r360_api.cu) and the kernels.
#pragma once
#include "r360_device.cuh"
#include "stitch_math.h"

#ifndef R360_PASS_THREADS
#define R360_PASS_THREADS 256                 // threads per CTA of k_pass
#endif
#ifndef R360_PASS_UNROLL
#define R360_PASS_UNROLL 2                    // main loop of k_pass unrolled x2 (stage parity becomes static): +1.7 % on B200
#endif
#ifndef R360_PASS_STAGES
#define R360_PASS_STAGES 2                    // depth of the shared-memory gather pipeline of k_pass
#endif
#ifndef R360_PASS_CTAS
#define R360_PASS_CTAS 2                      // resident CTAs per SM of k_pass (persistent grid = CTAS x SMs)
#endif
#ifndef R360_ERR_CTAS
#define R360_ERR_CTAS 3                       // resident CTAs per SM of the error-only pass (no 56 accumulator registers)
#endif

struct R360PassArgs {
    R360Level lv;
    r360_params params;
    float inv_std_photo;                // (float)(1./stdDevPhoto), RPI.h:2774
    float one;                          // 1.0f, opaque to the compiler (see f2add_sep)
    int items_per_pair, px_per_item;
    int dyn_permille;                   // share of the items (in 1/1000) handed out dynamically at the end of the launch
    int* work_counter;                  // device: next dynamic item (0 at launch; reset by the state-machine kernels)
    const int* n_active;                // device
    const int* active_list;             // device
    const R360Pair* pairs;              // device
    const float2* const* src_base;      // device: per pair, source pyramid {depth, gray}
    const float* const* trg_base;       // device: per pair, target texel pyramid
    R360Fx* acc;                        // device: per pair R360_ACC_STRIDE fixed-point sums (r360_fx_add / r360_fx_get)
    int* cnt;                           // device: per pair R360_ACC_INTS
};

// Pinhole intrinsics of one pyramid level, formed on the host exactly as the reference does
// (RPI.h:569-575: scaleFactor = 1/2^level as float, fx * scaleFactor, inv_fx = 1./fx narrowed to float).
struct R360PinLevel {
    float fx, fy, ox, oy, inv_fx, inv_fy;
};

// The 8-sensor rig (RegisterRGBD360::RegisterDensePhotoICP, RegisterRGBD360.h:344-520): extrinsics of the sensors,
// the per-pair frame tables and the intrinsics of one level in the two precisions the reference uses.
struct R360RigArgs {
    float Rt[8][16], Rt_inv[8][16];     // poseCamRobot and its inverse (r360_inverse4), column-major
    const float2* const* src8;          // device: per pair, the 8 source pyramids (frame2's sensors)
    const float* const* trg8;           // device: per pair, the 8 target texel pyramids (frame1's sensors)
    float fx, fy, ox, oy, inv_fx, inv_fy;            // calcPhotoICPError_robot: float (RPI.h:4915-4921)
    double dfx, dfy, dox, doy, dinv_fx, dinv_fy;     // calcHessianGradient_robot: double (RPI.h:5108-5114)
};

struct R360StitchArgs {
    R360StitchGeom g;
    float Rt_inv[8][16];                // inverse extrinsics of the 8 sensors, column-major (Calib360.h:122-131)
};

struct R360GnArgs {
    r360_params params;
    int n_pairs;
    R360Pair* pairs;
    R360Fx* acc;
    int* cnt;
    int* active_list;                   // pairs whose next pass is the fused one (error + normal equations)
    int* n_active;
    int* active_list_err;               // pairs whose next pass is error-only (speculation of k_gn_step)
    int* n_active_err;
    int speculate;                      // 0: every pass is the fused one
    float spec_margin;                  // error-only when the predicted RMS decrease < spec_margin * tol_residual (1: the plain rule)
    int* ticket;                        // device: block-completion counter (the last block compacts the active lists)
    int* work_counters;                 // device: the two dynamic-item counters of k_pass (fused, error-only): reset with the lists
    double lambda0;                     // > 0: the level's initial lambda (the rig driver: 0.001, RegisterRGBD360.h:391)
    int rig_faithful;                   // rig driver: 1 = new_error is evaluated at pose_estim, as upstream (RegisterRGBD360.h:462, 488)
    r360_iter_record* trace;            // device or nullptr
};

void r360_launch_level0(cudaStream_t st, const uint8_t* rgb, const uint16_t* depth_mm, const float* depth_m,
                        float2* const* dst, int n_frames, int n_px, int sm_count);
void r360_launch_down(cudaStream_t st, float2* const* pyr, long long off_src, long long off_dst, int rows,
                      int cols, float min_d, float max_d, int n_frames, int sm_count);
void r360_launch_texel(cudaStream_t st, float2* const* pyr, float* const* trg, long long off, int rows, int cols,
                       int n_sensors, int n_frames, int sm_count);
// fused level 0 + level-0 texels + level 1; l0_dst[f] / texel_dst[f] may be null (role), l1_dst[f] points at level 1
void r360_launch_pyr_head(cudaStream_t st, const uint8_t* rgb, const uint16_t* depth_mm, const float* depth_m,
                          float2* const* l0_dst, float2* const* l1_dst, float* const* texel_dst, int rows, int cols,
                          float min_d, float max_d, int n_sensors, int n_frames);
void r360_launch_pyr_mid(cudaStream_t st, float2* const* pyr, float* const* tex, long long off, long long off_next, int rows, int cols,
                         float min_d, float max_d, int n_sensors, int n_frames);
cudaError_t r360_pass_init();
void r360_launch_pass(cudaStream_t st, const R360PassArgs& a, int grid, bool with_h = true);
// occlusion variants (r360_occ.cu): head / next / dinv hold n_pairs * lv.n entries each
cudaError_t r360_occ_init();
void r360_launch_occ_pass(cudaStream_t st, const R360PassArgs& a, int n_pairs, int* head, int* next, float* dinv,
                          int sm_count);
// pinhole registration (r360_pinhole.cuh)
void r360_launch_pin_eval(cudaStream_t st, const R360PassArgs& a, const R360PinLevel& pl, int n_pairs, int sm_count, bool packed = true);
void r360_launch_gn_step_pin(cudaStream_t st, const R360GnArgs& g, int level);
// the 8-sensor rig (r360_pinhole.cuh)
void r360_launch_rig_eval(cudaStream_t st, const R360PassArgs& a, const R360RigArgs& rig, int n_pairs, int sm_count);
void r360_launch_gn_step_rig(cudaStream_t st, const R360GnArgs& g, int level);
void r360_launch_warp_dump(cudaStream_t st, const R360PassArgs& a, int pair, int32_t* r_idx, int32_t* c_idx,
                           uint8_t* vp, uint8_t* vd, int sm_count);
void r360_launch_index_stats(cudaStream_t st, const R360PassArgs& a, int pair, unsigned long long* out, int sm_count);
void r360_launch_stitch(cudaStream_t st, const R360StitchArgs& a, const uint8_t* sensor_rgb, const uint16_t* sensor_depth,
                        uint8_t* rgb, uint16_t* depth_mm, int n_frames, int sm_count);
void r360_launch_pairs_init(cudaStream_t st, const R360GnArgs& g, const int32_t* src_idx, const int32_t* trg_idx,
                            const float* init_pose);
void r360_launch_level_begin(cudaStream_t st, const R360GnArgs& g, int level);
void r360_launch_gn_step(cudaStream_t st, const R360GnArgs& g, int level);
void r360_launch_finalize(cudaStream_t st, const R360GnArgs& g, r360_result* out, int rows, int cols, int pair_id0);
void r360_launch_synth(cudaStream_t st, int kind, int first_id, int rows, int cols, const float* cams, int n_frames,
                       uint8_t* rgb, uint16_t* depth_mm, int sm_count);
