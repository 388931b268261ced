/* r360.h -- C ABI of the B200 spherical dense-registration path.
 *
 * Drop-in boundary for ONE path of EduFdez/rgbd360: RegisterPhotoICP's spherical
 * photometric + depth registration (include/RegisterPhotoICP.h, "RPI.h" below).
 * The reference has no FFI layer (RegisterPhotoICP is a header-only C++ class), so each
 * entry point below names the class method / loop it replaces.  Plain pointers and
 * sizes only; all pointers are HOST pointers unless the name ends in _dev.
 * A C++ mirror of the class (same method names) lives in include/RegisterPhotoICP_b200.hpp.
 *
 * Threading: one r360_ctx per GPU, not thread-safe per ctx (same contract as the
 * reference class, which mutates members in every call).
 * Errors: every call returns 0 on success, a negative R360_E_* code otherwise, and
 * r360_last_error(ctx) returns text.  There is NO CPU fallback: without a CUDA device
 * r360_create fails with R360_E_CUDA.
 */
#ifndef R360_H
#define R360_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define R360_MAX_LEVELS 8

enum { R360_OK = 0, R360_E_ARG = -1, R360_E_CUDA = -2, R360_E_STATE = -3, R360_E_NOMEM = -4 };

/* costFuncType, RPI.h:195 */
enum { R360_PHOTO_CONSISTENCY = 0, R360_DEPTH_CONSISTENCY = 1, R360_PHOTO_DEPTH = 2 };

/* per-pair status (the reference prints "ILL-POSED" and returns, RPI.h:4682-4690) */
enum { R360_PAIR_OK = 0, R360_PAIR_ILL_POSED = 1 };

/* which registration of the class a context runs */
enum { R360_SPHERE = 0, R360_PINHOLE = 1 };

/* frame roles for r360_set_frames */
enum { R360_ROLE_SOURCE = 1, R360_ROLE_TARGET = 2, R360_ROLE_BOTH = 3 };

/* Setters of RegisterPhotoICP (RPI.h:224-269) + the constants hard-coded in
 * alignFrames360 (RPI.h:4589-4596).  r360_default_params() fills the ctor defaults
 * (RPI.h:201-221). */
typedef struct r360_params {
    int32_t n_levels;        /* setNumPyr            (default 4)                        */
    float   min_depth;       /* setMinDepth          (0.3)                              */
    float   max_depth;       /* setMaxDepth          (6.0)                              */
    float   std_photo;       /* setGrayVariance      (6/255; it sets stdDevPhoto)       */
    float   std_depth;       /* setDepthVariance     (0.2)                              */
    float   thres_sal_int;   /* thresSaliencyIntensity (0.01)                           */
    float   thres_sal_depth; /* thresSaliencyDepth     (0.01)                           */
    int32_t max_iters;       /* maxIters = 10        RPI.h:4593                         */
    double  tol_residual;    /* 1e-3                 RPI.h:4594                         */
    double  tol_update;      /* 1e-4                 RPI.h:4595                         */
    int32_t method;          /* costFuncType; callers use R360_PHOTO_DEPTH              */
    int32_t occlusion;       /* alignFrames360's `occlusion` argument (RPI.h:4519): 0 regular,
                                1 = *_sphereOcc1 (z-buffer), 2 = *_sphereOcc2 (0.3 m outlier gate +
                                z-buffer), with the reference's single-thread (source-order) semantics */
    int32_t n_sensors_mask;  /* 8: zero the 2-px sensor-joint gradient columns
                                (RPI.h:4537-4549) when the target pyramid is built; 0: no mask */
    int32_t projection;      /* R360_SPHERE (alignFrames360 and friends) or R360_PINHOLE (alignFrames,
                                errorPhotoICP, calcHessGrad: RPI.h:4254, 560, 776; needs r360_set_camera) */
} r360_params;

/* What the getters and public fields of the class expose after alignFrames360
 * (RPI.h:273-288, 179-189), one record per pair.  POD, fixed size: it is the unit
 * that is all-gathered across GPUs. */
typedef struct r360_result {
    float   pose[16];        /* getOptimalPose(), column-major 4x4 (Eigen layout)       */
    float   hessian[36];     /* getHessian(): of the LAST calcHessGrad_sphere call      */
    float   gradient[6];     /* getGradient()                                           */
    float   sso;             /* SSO = numVisiblePixels / imgSize   RPI.h:3226           */
    int32_t n_visible;       /* numVisiblePixels of that call                           */
    double  final_error;     /* `error` (RMS) at the accepted pose of the last level run */
    double  final_err2;      /* its sum of squared weighted residuals (occlusion 1/2:
                                PhotoResidual + DepthResidual)                          */
    int32_t final_n_valid;   /* its numValidPts (occlusion 1: nValidPhotoPts + nValidDepthPts,
                                occlusion 2: nValidDepthPts)                            */
    int32_t status;          /* R360_PAIR_*                                             */
    int32_t iters[R360_MAX_LEVELS];   /* num_iterations[level] (accepted steps)         */
    int32_t passes[R360_MAX_LEVELS];  /* fused pixel passes executed per level          */
    int32_t pair_id;         /* index of the pair in the call                           */
    int32_t reserved;
} r360_result;

/* Optional per-iteration trace (parity hook): one record per evaluated pose. */
typedef struct r360_iter_record {
    double  err2;            /* sum of squared weighted residuals at `pose`
                                (occlusion 1/2: PhotoResidual)                          */
    int32_t n_valid;         /* numValidPts (occlusion 1: nValidPhotoPts, 2: nValidDepthPts) */
    int32_t n_visible;
    int32_t level;
    int32_t it;              /* accepted steps so far when this pose was evaluated      */
    int32_t accepted;        /* 1: became pose_estim (the initial evaluation counts)    */
    int32_t used;            /* 1: record holds data                                    */
    float   pose[16];
    float   hessian[21];     /* upper triangle, row-major (h11,h12,..,h66)              */
    float   gradient[6];
    float   pad;
    double  err2_depth;      /* occlusion 1/2 and pinhole: DepthResidual (RPI.h:3348, 3725, 563), else 0 */
    int32_t n_valid_depth;   /* occlusion 1/2 and pinhole: nValidDepthPts, else 0       */
    int32_t reserved;
} r360_iter_record;

typedef struct r360_ctx r360_ctx;

void r360_default_params(r360_params* p);
/* The constants hard-coded in the pinhole alignFrames (RPI.h:4304-4309: lambda 0.01, step 10 -- fixed in the
 * state machine -- and tol_residual 1e-4), projection = R360_PINHOLE, no sensor-joint mask (that mask lives
 * in alignFrames360, RPI.h:4537). */
void r360_default_params_pinhole(r360_params* p);
const char* r360_last_error(const r360_ctx* ctx);   /* ctx may be NULL: last create error */

/* Replaces constructing RegisterPhotoICP instances (RPI.h:201) for a whole batch:
 * `max_frames` frame slots of rows x cols, up to `max_pairs` pairs per call. */
int r360_create(r360_ctx** ctx, int device, int rows, int cols, int max_frames, int max_pairs,
                const r360_params* params);
void r360_destroy(r360_ctx* ctx);

/* setSourceFrame / setTargetFrame (RPI.h:480-516) for frames [first, first+n):
 * gray conversion, gray / depth pyramids and, for target roles, the gradient pyramids
 * with the sensor-joint mask.  rgb: n x rows x cols x 3 u8 (channel 0 is taken as R,
 * as the reference does); depth_mm: n x rows x cols u16 millimetres.
 * roles: n entries of R360_ROLE_* or NULL (= BOTH). */
int r360_set_frames(r360_ctx* ctx, int first, int n, const uint8_t* rgb, const uint16_t* depth_mm,
                    const uint8_t* roles);
/* Same, inputs already in device memory (the HBM-resident form the throughput metric uses). */
int r360_set_frames_dev(r360_ctx* ctx, int first, int n, const uint8_t* rgb_dev,
                        const uint16_t* depth_mm_dev, const uint8_t* roles);
/* setSourceFrame/setTargetFrame with a CV_32FC1 depth image in metres (RPI.h:316-319). */
int r360_set_frames_f32(r360_ctx* ctx, int first, int n, const uint8_t* rgb, const float* depth_m,
                        const uint8_t* roles);

/* alignFrames360(pose_guess, method, occlusion = 0) (RPI.h:4519-4784) for n_pairs
 * independent (source, target) pairs.  init_pose: n_pairs x 16 floats column-major, or
 * NULL for Identity.  trace: NULL or n_pairs * n_levels * (max_iters + 2) records. */
int r360_register_pairs(r360_ctx* ctx, int n_pairs, const int32_t* src_idx, const int32_t* trg_idx,
                        const float* init_pose, r360_result* out, r360_iter_record* trace);

/* The per-pair sequence setTargetFrame + setSourceFrame + alignFrames360 (RPI.h:498-516, 480-495,
 * 4519; call sites Registration/OdometryRGBD360.cpp:189-193, include/LoopClosure360.h:306-321)
 * for n_pairs pairs whose frames are in HOST memory, as ONE pipelined call: frame 2p of
 * rgb / depth_mm is the target and frame 2p+1 the source of pair p (they occupy frame slots
 * 2p and 2p+1 afterwards; needs max_frames >= 2 n_pairs).  Uploads, pyramid builds and batched
 * registrations overlap on the device; the host blocks once.  Pinned host buffers recommended. */
int r360_register_host_pairs(r360_ctx* ctx, int n_pairs, const uint8_t* rgb, const uint16_t* depth_mm,
                             const float* init_pose, r360_result* out);

/* errorPhotoICP_sphere(level, pose, method) (RPI.h:2545-2739): returns the two sums the
 * RMS is formed from.  Contexts created with occlusion 1/2 return PhotoResidual + DepthResidual
 * and the sum of the counters; use r360_eval_error_occ for the separate terms. */
int r360_eval_error(r360_ctx* ctx, int src, int trg, int level, const float pose[16],
                    double* err2, int32_t* n_valid);
/* errorPhotoICP_sphereOcc1 / errorPhotoICP_sphereOcc2(level, pose, method) (RPI.h:3232-3369,
 * 3720-3858) of a context created with params.occlusion = 1 / 2: PhotoResidual, DepthResidual,
 * nValidPhotoPts (0 for occlusion 2, which does not count it) and nValidDepthPts; `error` = the
 * function's return value avPhotoResidual + avDepthResidual.  Any output pointer may be NULL. */
int r360_eval_error_occ(r360_ctx* ctx, int src, int trg, int level, const float pose[16],
                        double* photo_residual, double* depth_residual, int32_t* n_valid_photo,
                        int32_t* n_valid_depth, double* error);
/* calcHessGrad_sphere(level, pose, method) (RPI.h:2745-3228): H column-major 6x6.  With
 * params.occlusion = 1 / 2: calcHessGrad_sphereOcc1 / Occ2 (RPI.h:3373-3716, 3861-4249). */
int r360_eval_hessgrad(r360_ctx* ctx, int src, int trg, int level, const float pose[16],
                       float H[36], float g[6], int32_t* n_visible);

/* ---- pinhole registration, contexts created with projection = R360_PINHOLE ----
 * setCameraMatrix (RPI.h:254): fx, fy, ox, oy of the level-0 images; the levels scale them by 2^-level
 * (RPI.h:569-573).  r360_register_pairs then runs alignFrames(guess, method, occlusion = 0) (RPI.h:4254-4512:
 * Levenberg-Marquardt with one damped retry, full SE(3) exponential); trace needs
 * n_pairs * n_levels * (2 * max_iters + 2) records.  r360_eval_hessgrad = calcHessGrad (RPI.h:776). */
int r360_set_camera(r360_ctx* ctx, float fx, float fy, float ox, float oy);
/* errorPhotoICP(level, pose, method) (RPI.h:560-775): PhotoResidual, DepthResidual, nValidPhotoPts,
 * nValidDepthPts and the returned avResidual = (float)(sqrt(Photo / nValidDepthPts) + sqrt(Depth / nValidDepthPts)). */
int r360_eval_error_pinhole(r360_ctx* ctx, int src, int trg, int level, const float pose[16],
                            double* photo_residual, double* depth_residual, int32_t* n_valid_photo,
                            int32_t* n_valid_depth, double* error);

/* ---- the 8-sensor rig: RegisterRGBD360::RegisterDensePhotoICP (include/RegisterRGBD360.h:344-520) ----
 * The dense registration of two Frame360 over their 8 pinhole sensor images: per sensor calcPhotoICPError_robot
 * (RPI.h:4905-5092) and calcHessianGradient_robot (RPI.h:5100-5407) with the sensor's extrinsics, errors / Hessians /
 * gradients summed over the rig, one 6-DoF Levenberg-Marquardt loop on the ROBOT pose (lambda 0.001, step 10,
 * tol_residual 0.1, tol_update 1e-6, 10 iterations: RegisterRGBD360.h:391-398, fixed here whatever the ctx params say).
 * Contexts created with projection = R360_PINHOLE, method = R360_PHOTO_CONSISTENCY (the driver's default; upstream the
 * Hessian's depth row reads a matrix that is never assigned, RPI.h:5372-5374, so the other methods are undefined there
 * and refused here) and r360_set_camera(f, f, w/2 - 0.5, h/2 - 0.5), f = 525 w / 640 (RegisterRGBD360.h:361-369).
 * A rig frame occupies 8 consecutive frame slots (sensor s at first + s): src_first / trg_first name sensor 0 of
 * frame2 (source) / frame1 (target).  Rt: the 8 sensor poses calib->Rt_ (column-major 4x4 each), shared by all pairs.
 * faithful_new_error != 0: `new_error` is evaluated at pose_estim, as upstream does (RegisterRGBD360.h:462, 488) --
 * diff_error is then exactly 0, no step is ever taken, and the call returns the initial guess with the rig's summed
 * Hessian at that guess (what upstream returns whenever its OpenMP reduction happens to sum in the same order twice);
 * 0: the candidate pose is evaluated (the evident intent).  out[p].pose = rigidTransf, .hessian = informationM,
 * .status = R360_PAIR_ILL_POSED where upstream returns false, .final_error = the summed squared error. */
int r360_register_rig_pairs(r360_ctx* ctx, int n_pairs, const int32_t* src_first, const int32_t* trg_first,
                            const float* Rt, const float* init_pose, int faithful_new_error, r360_result* out);
/* One evaluation of the rig at `pose`: error2 = sum over the sensors of calcPhotoICPError_robot, H (column-major 6x6)
 * and g = the sums of calcHessianGradient_robot's hessian / gradient.  Any output pointer may be NULL. */
int r360_eval_rig(r360_ctx* ctx, int src_first, int trg_first, int level, const float pose[16], const float* Rt,
                  double* error2, float H[36], float g[6], int32_t* n_visible, int32_t* n_error_terms);

/* Parity hooks (a1-a5 planes, warp index maps).  Any output pointer may be NULL. */
int r360_dump_level(r360_ctx* ctx, int frame, int level, float* gray, float* depth,
                    float* gray_gx, float* gray_gy, float* depth_gx, float* depth_gy);
/* Source-role planes ({depth, gray} pyramid read by the warp) of a frame. */
int r360_dump_source_level(r360_ctx* ctx, int frame, int level, float* gray, float* depth);
int r360_dump_warp(r360_ctx* ctx, int src, int trg, int level, const float pose[16],
                   int32_t* r_idx, int32_t* c_idx, uint8_t* valid_photo, uint8_t* valid_depth);

/* The warp kernels evaluate the pinned index sequence on two pixels at once with packed fp32x2
 * instructions and send pixels outside the range where that is exact (denormal/huge operands,
 * exact .5 ties) to the scalar pinned code.  r360_index_stats cross-checks the two over every
 * valid source pixel of a pair: out[0] valid pixels, out[1] pixels sent to the scalar code,
 * out[2] pixels kept by the packed code whose (r', c') differ from the scalar result (must be 0). */
int r360_index_stats(r360_ctx* ctx, int src, int trg, int level, const float pose[16], uint64_t out[3]);

/* Synthetic sphere frames of SURVEY 8(d) rendered straight into device buffers
 * (convex box room, procedural texture); frame ids and the scene kind select poses.
 * kind 0: odometry / batch trajectory, 1: loop-closure keyframes. */
int r360_synth_frames_dev(r360_ctx* ctx, int kind, int first_id, int n, uint8_t* rgb_dev,
                          uint16_t* depth_mm_dev);
int r360_synth_frames(r360_ctx* ctx, int kind, int first_id, int n, uint8_t* rgb, uint16_t* depth_mm);
/* Ground-truth pose T_trg<-src (column-major 4x4 double) between two synthetic frames. */
void r360_synth_gt_pose(int kind, int src_id, int trg_id, double T[16]);

/* ---- Frame360 ingest: the step right before the path --------------------------
 * Calib360 (include/Calib360.h:69-131): pinhole intrinsics shared by the 8 sensors and the INVERSE
 * of each sensor's extrinsic matrix (Rt_inv[s] = Rt_[s].inverse(), column-major 4x4). */
typedef struct r360_rig {
    int32_t sensor_rows, sensor_cols;   /* 240, 320 (QVGA, Calib360.h:63-65)                      */
    float   fx, fy, cx, cy;             /* 262.5, 262.5, 159.5, 119.5 (Calib360.h:75-77)           */
    float   Rt_inv[8][16];
} r360_rig;
void r360_default_rig(r360_rig* rig);   /* QVGA intrinsics, identity extrinsics */

/* Frame360::loadFrame (include/Frame360.h:231-266): parses a Frame360 .bin (boost binary_oarchive of
 * 16 cv::Mat records, third_party/cvSerialization/cvmat_serialization.h:22-54, alternating RGB
 * CV_8UC3 and depth CV_16UC1 for sensors 0..7).  rgb: 8 x rows x cols x 3, depth_mm: 8 x rows x
 * cols; either may be NULL to query the size only.  Host-only, no ctx needed. */
int r360_frame360_parse(const uint8_t* bytes, size_t n_bytes, int32_t* sensor_rows, int32_t* sensor_cols,
                        uint8_t* rgb, size_t rgb_capacity, uint16_t* depth_mm, size_t depth_capacity);

/* Frame360::stitchSphericalImage (include/Frame360.h:386-405, 1099-1148) on the device for n
 * frames, followed by setSourceFrame / setTargetFrame of the stitched spheres into frame slots
 * [first, first+n) (as r360_set_frames).  sensor_rgb: n x 8 x sensor_rows x sensor_cols x 3,
 * sensor_depth_mm: n x 8 x sensor_rows x sensor_cols (z-depth, millimetres).  The ctx must have been
 * created with cols = 8 * sensor_rows and rows = (int)(cols * 0.5 * 60 / 180).
 * sphere_rgb / sphere_depth_mm: optional host outputs (n x rows x cols [x 3]) of what the reference
 * keeps as Frame360::sphereRGB / sphereDepth. */
int r360_stitch_frames(r360_ctx* ctx, const r360_rig* rig, int first, int n, const uint8_t* sensor_rgb,
                       const uint16_t* sensor_depth_mm, const uint8_t* roles, uint8_t* sphere_rgb,
                       uint16_t* sphere_depth_mm);

/* Multi-GPU: the path shards by pairs, one r360_ctx (and one NCCL rank) per GPU; the only exchange is
 * ONE all-gather of the fixed-size result records over NVLink / NVSwitch.  `nccl_comm` is the caller's ncclComm_t for
 * this ctx's device.  NCCL is resolved at run time -- the ncclAllGather already loaded in the host process, else
 * dlopen("libnccl.so.2") -- so this library carries no NCCL dependency; R360_E_STATE when none is found.
 * local: n_local records (host); all: n_ranks * n_local records (host), rank-major.  Every rank passes the same
 * n_local (pad a ragged last shard with records whose pair_id is -1, as rgbd360_b200/shard.py does).  Collective: every rank of the
 * communicator must call it; returns after the gathered records are in `all`.
 * Replaces nothing upstream (the reference registers its pairs one after another on one host, LoopClosure360.h:291-321,
 * OdometryRGBD360.cpp:185-193); it returns what those loops accumulate. */
int r360_allgather_results(r360_ctx* ctx, void* nccl_comm, const r360_result* local, int n_local, int n_ranks,
                           r360_result* all);

/* Pinned host staging memory for the frames a caller hands to r360_set_frames / r360_register_host_pairs
 * (the reference's callers hold them as cv::Mat, OdometryRGBD360.cpp:185-193), placed on the NUMA node the
 * ctx's GPU hangs off: with one rank per GPU all uploading at once, staging memory on the wrong socket
 * halves the host-to-device rate.  The node comes from the GPU's PCI entry in sysfs; the pages are bound
 * with mbind(2) when the process may, else first-touched by this thread pinned to the node's CPUs, then
 * page-locked with cudaHostRegister.  *numa_node: the node the first page actually sits on afterwards
 * (get_mempolicy), or -1 when the platform exposes no NUMA topology for the device (single node, VM) -- then
 * the memory is plain cudaHostAlloc memory and r360_last_error(ctx) says why.  Free with r360_host_free. */
int r360_host_alloc(r360_ctx* ctx, size_t bytes, void** ptr, int32_t* numa_node);
int r360_host_free(r360_ctx* ctx, void* ptr);

/* Device memory / stream plumbing for callers that keep data on the GPU. */
int r360_device_alloc(r360_ctx* ctx, size_t bytes, void** ptr_dev);
int r360_device_free(r360_ctx* ctx, void* ptr_dev);
int r360_synchronize(r360_ctx* ctx);
/* Milliseconds spent in the last r360_register_pairs / r360_set_frames* call measured with
 * CUDA events on the ctx stream, and kernels launched by the ctx since creation. */
float r360_last_device_ms(const r360_ctx* ctx);
int64_t r360_kernel_launches(const r360_ctx* ctx);
/* Average device time (ms) and launch count of the fused warp/residual/normal-equation
 * kernel inside the last r360_register_pairs call, and the algorithmic bytes it moved. */
int r360_last_pass_stats(const r360_ctx* ctx, float* total_ms, int32_t* launches, double* alg_bytes);

int r360_version(void);

#ifdef __cplusplus
}
#endif
#endif /* R360_H */
