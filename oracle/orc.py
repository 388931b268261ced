"""ctypes binding of the CPU oracle (oracle/rpi_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never by the product package rgbd360_b200.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "librpi_oracle.so")

R360_MAX_LEVELS = 8
PHOTO, DEPTH, PHOTO_DEPTH = 0, 1, 2
MATH_PINNED, MATH_LIBM = 0, 1
ACC_FAITHFUL, ACC_STABLE = 0, 1


class Params(C.Structure):
    _fields_ = [
        ("n_levels", C.c_int32), ("min_depth", C.c_float), ("max_depth", C.c_float),
        ("std_photo", C.c_float), ("std_depth", C.c_float), ("thres_sal_int", C.c_float),
        ("thres_sal_depth", C.c_float), ("max_iters", C.c_int32), ("tol_residual", C.c_double),
        ("tol_update", C.c_double), ("method", C.c_int32), ("occlusion", C.c_int32),
        ("n_sensors_mask", C.c_int32), ("projection", C.c_int32),
    ]


class Result(C.Structure):
    _fields_ = [
        ("pose", C.c_float * 16), ("hessian", C.c_float * 36), ("gradient", C.c_float * 6),
        ("sso", C.c_float), ("n_visible", C.c_int32), ("final_error", C.c_double),
        ("final_err2", C.c_double), ("final_n_valid", C.c_int32), ("status", C.c_int32),
        ("iters", C.c_int32 * R360_MAX_LEVELS), ("passes", C.c_int32 * R360_MAX_LEVELS),
        ("pair_id", C.c_int32), ("reserved", C.c_int32),
    ]


class IterRecord(C.Structure):
    _fields_ = [
        ("err2", C.c_double), ("n_valid", C.c_int32), ("n_visible", C.c_int32),
        ("level", C.c_int32), ("it", C.c_int32), ("accepted", C.c_int32), ("used", C.c_int32),
        ("pose", C.c_float * 16), ("hessian", C.c_float * 21), ("gradient", C.c_float * 6),
        ("pad", C.c_float), ("err2_depth", C.c_double), ("n_valid_depth", C.c_int32), ("reserved", C.c_int32),
    ]


def default_params(n_levels=4, method=PHOTO_DEPTH, std_photo=None, n_sensors_mask=8, occlusion=0):
    """Constructor defaults of RegisterPhotoICP (RPI.h:201-221) + alignFrames360 constants."""
    p = Params()
    p.n_levels = n_levels
    p.min_depth = 0.3
    p.max_depth = 6.0
    p.std_photo = np.float32(6.0 / 255) if std_photo is None else np.float32(std_photo)
    p.std_depth = 0.2
    p.thres_sal_int = 0.01
    p.thres_sal_depth = 0.01
    p.max_iters = 10
    p.tol_residual = 1e-3
    p.tol_update = 1e-4
    p.method = method
    p.occlusion = occlusion
    p.n_sensors_mask = n_sensors_mask
    return p


def build(force=False):
    """Compile the oracle with oracle/Makefile (g++ only)."""
    if force or not os.path.exists(_SO) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_SO)
        for f in ("rpi_oracle.cpp", "../rgbd360_b200/csrc/sphere_math.h",
                  "../rgbd360_b200/csrc/gn_math.h", "../rgbd360_b200/csrc/synth.h",
                  "../rgbd360_b200/csrc/stitch_math.h", "../include/r360.h")
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.orc_frame_build.restype = C.c_void_p
        L.orc_frame_build.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                      C.POINTER(Params), C.c_int]
        L.orc_frame_free.argtypes = [C.c_void_p]
        L.orc_frame_level.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 6
        L.orc_lut.argtypes = [C.c_void_p, C.c_int, C.POINTER(Params), C.c_void_p]
        L.orc_error.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(Params),
                                C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.orc_error_occ.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(Params),
                                    C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        L.orc_hessgrad.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(Params),
                                   C.c_int] + [C.c_void_p] * 5
        L.orc_warp.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(Params)] + [C.c_void_p] * 4
        L.orc_align.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Params), C.c_int,
                                C.POINTER(Result), C.c_void_p, C.c_int]
        L.orc_synth_frame.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_error_pinhole.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(Params), C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        L.orc_hessgrad_pinhole.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(Params), C.c_void_p,
                                           C.c_int] + [C.c_void_p] * 5
        L.orc_align_pinhole.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_int,
                                        C.POINTER(Result), C.c_void_p, C.c_int]
        L.orc_synth_pinhole_frame.argtypes = [C.c_int] * 4 + [C.c_float] * 4 + [C.c_void_p, C.c_void_p]
        L.orc_synth_pinhole_frame_rt.argtypes = [C.c_int] * 4 + [C.c_float] * 4 + [C.c_void_p] * 3
        L.orc_error_robot.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(Params), C.c_void_p,
                                      C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.orc_hessgrad_robot.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(Params), C.c_void_p] + [C.c_void_p] * 5
        L.orc_align_rig.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_int, C.c_int,
                                    C.POINTER(Result)]
        L.orc_inverse4.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_synth_gt_pose.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_stitch.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float] + [C.c_void_p] * 5
        L.orc_pinned_vec.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_pinned_sincos.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_rank6.argtypes = [C.c_void_p]
        L.orc_inverse6.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_solve_update.argtypes = [C.c_void_p] * 3
        L.orc_pseudo_exp.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_mat4_mul.argtypes = [C.c_void_p] * 3
        _lib = L
    return _lib


def set_math(mode):
    lib().orc_set_math(int(mode))


def omp_threads():
    return int(lib().orc_omp_threads())


def use_all_cores():
    """OpenMP threads = the cores this process may run on (torchrun exports OMP_NUM_THREADS=1)."""
    import os
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().orc_set_threads(n)
    return n


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def pose_arg(pose):
    """4x4 (row-major numpy view of a matrix) -> 16 floats column-major (Eigen layout)."""
    if pose is None:
        pose = np.eye(4)
    return np.ascontiguousarray(np.asarray(pose, dtype=np.float32).reshape(4, 4).T).reshape(16)


def pose_from(buf):
    return np.array(buf, dtype=np.float32).reshape(4, 4).T.copy()


class Frame:
    """One side of a pair: what setSourceFrame / setTargetFrame build (RPI.h:480-516)."""

    def __init__(self, rgb, depth, params, with_grad=True):
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
        self.rows, self.cols = rgb.shape[:2]
        self.params = params
        self.with_grad = with_grad
        if depth.dtype == np.uint16:
            d = np.ascontiguousarray(depth)
            self.h = lib().orc_frame_build(_ptr(rgb), _ptr(d), None, self.rows, self.cols, C.byref(params), int(with_grad))
        else:
            d = np.ascontiguousarray(depth, dtype=np.float32)
            self.h = lib().orc_frame_build(_ptr(rgb), None, _ptr(d), self.rows, self.cols, C.byref(params), int(with_grad))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_frame_free(self.h)
            self.h = None

    def level(self, level):
        r, c = self.rows >> level, self.cols >> level
        names = ["gray", "depth"] + (["ggx", "ggy", "dgx", "dgy"] if self.with_grad else [])
        out = {n: np.zeros((r, c), np.float32) for n in names}
        args = [_ptr(out[n]) if n in out else None for n in ["gray", "depth", "ggx", "ggy", "dgx", "dgy"]]
        lib().orc_frame_level(self.h, level, *args)
        return out


def lut(src, level, params):
    r, c = src.rows >> level, src.cols >> level
    out = np.zeros((r * c, 3), np.float32)
    lib().orc_lut(src.h, level, C.byref(params), _ptr(out))
    return out


def error(src, trg, level, pose, params):
    e2, n = C.c_double(), C.c_int()
    T = pose_arg(pose)
    lib().orc_error(src.h, trg.h, level, _ptr(T), C.byref(params), C.byref(e2), C.byref(n))
    return e2.value, n.value


def error_occ(src, trg, level, pose, params):
    """errorPhotoICP_sphereOcc1 / Occ2 (params.occlusion) -> dict(photo, depth, n_photo, n_depth, error)."""
    r2 = np.zeros(2, np.float64); cnt = np.zeros(2, np.int32); e = C.c_double()
    T = pose_arg(pose)
    lib().orc_error_occ(src.h, trg.h, level, _ptr(T), C.byref(params), _ptr(r2), _ptr(cnt), C.byref(e))
    return dict(photo=float(r2[0]), depth=float(r2[1]), n_photo=int(cnt[0]), n_depth=int(cnt[1]), error=e.value)


def hessgrad(src, trg, level, pose, params, accum=ACC_STABLE):
    H = np.zeros(36, np.float32); g = np.zeros(6, np.float32)
    Hd = np.zeros(21, np.float64); gd = np.zeros(6, np.float64); cnt = np.zeros(3, np.int32)
    T = pose_arg(pose)
    lib().orc_hessgrad(src.h, trg.h, level, _ptr(T), C.byref(params), accum, _ptr(H), _ptr(g), _ptr(Hd), _ptr(gd), _ptr(cnt))
    return dict(H=H.reshape(6, 6), g=g, Hd=Hd, gd=gd, n_visible=int(cnt[0]), n_photo=int(cnt[1]), n_depth=int(cnt[2]))


def warp(src, trg, level, pose, params):
    n = (src.rows >> level) * (src.cols >> level)
    ri = np.zeros(n, np.int32); ci = np.zeros(n, np.int32)
    vp = np.zeros(n, np.uint8); vd = np.zeros(n, np.uint8)
    T = pose_arg(pose)
    lib().orc_warp(src.h, trg.h, level, _ptr(T), C.byref(params), _ptr(ri), _ptr(ci), _ptr(vp), _ptr(vd))
    return ri, ci, vp, vd


def align(src, trg, guess, params, accum=ACC_STABLE, trace=False):
    res = Result()
    T = pose_arg(guess)
    cap = params.n_levels * (params.max_iters + 2)
    tr = (IterRecord * cap)() if trace else None
    lib().orc_align(src.h, trg.h, _ptr(T), C.byref(params), accum, C.byref(res),
                    C.cast(tr, C.c_void_p) if trace else None, cap if trace else 0)
    return (res, tr) if trace else res


def synth_frame(kind, fid, rows, cols):
    rgb = np.zeros((rows, cols, 3), np.uint8)
    d = np.zeros((rows, cols), np.uint16)
    lib().orc_synth_frame(kind, fid, rows, cols, _ptr(rgb), _ptr(d))
    return rgb, d


def pinhole_params(n_levels=4, method=PHOTO_DEPTH, std_photo=None):
    """Defaults of the pinhole alignFrames (RPI.h:4304-4309): tol_residual 1e-4, no sensor-joint mask."""
    p = default_params(n_levels=n_levels, method=method, std_photo=std_photo, n_sensors_mask=0)
    p.tol_residual = 1e-4
    p.projection = 1
    return p


def _cam(cam):
    return np.ascontiguousarray(cam, np.float32)


def error_pinhole(src, trg, level, pose, params, cam):
    r2 = np.zeros(2, np.float64); cnt = np.zeros(2, np.int32); e = C.c_double()
    T = pose_arg(pose); K = _cam(cam)
    lib().orc_error_pinhole(src.h, trg.h, level, _ptr(T), C.byref(params), _ptr(K), _ptr(r2), _ptr(cnt), C.byref(e))
    return dict(photo=float(r2[0]), depth=float(r2[1]), n_photo=int(cnt[0]), n_depth=int(cnt[1]), error=e.value)


def hessgrad_pinhole(src, trg, level, pose, params, cam, accum=ACC_STABLE):
    H = np.zeros(36, np.float32); g = np.zeros(6, np.float32)
    Hd = np.zeros(21, np.float64); gd = np.zeros(6, np.float64); cnt = np.zeros(3, np.int32)
    T = pose_arg(pose); K = _cam(cam)
    lib().orc_hessgrad_pinhole(src.h, trg.h, level, _ptr(T), C.byref(params), _ptr(K), accum, _ptr(H), _ptr(g), _ptr(Hd),
                               _ptr(gd), _ptr(cnt))
    return dict(H=H.reshape(6, 6), g=g, Hd=Hd, gd=gd, n_visible=int(cnt[0]), n_photo=int(cnt[1]), n_depth=int(cnt[2]))


def align_pinhole(src, trg, guess, params, cam, accum=ACC_STABLE, trace=False):
    res = Result()
    T = pose_arg(guess); K = _cam(cam)
    cap = params.n_levels * (2 * params.max_iters + 2)
    tr = (IterRecord * cap)() if trace else None
    lib().orc_align_pinhole(src.h, trg.h, _ptr(T), C.byref(params), _ptr(K), accum, C.byref(res),
                            C.cast(tr, C.c_void_p) if trace else None, cap if trace else 0)
    return (res, tr) if trace else res


def synth_pinhole_frame(kind, fid, rows, cols, fx, fy, ox, oy):
    """Pinhole view of the synthetic room from the pose of frame `fid` (depth = z in mm)."""
    rgb = np.zeros((rows, cols, 3), np.uint8)
    d = np.zeros((rows, cols), np.uint16)
    lib().orc_synth_pinhole_frame(kind, fid, rows, cols, fx, fy, ox, oy, _ptr(rgb), _ptr(d))
    return rgb, d


# ---- the 8-sensor rig: RegisterRGBD360::RegisterDensePhotoICP and the *_robot functions it sums
def rig_camera(rows, cols):
    """camIntrinsicMat of RegisterDensePhotoICP (RegisterRGBD360.h:361-369): (fx, fy, ox, oy) as float32."""
    f32 = np.float32
    res_factor_VGA = f32(f32(cols) / f32(640.0))
    focal = f32(f32(525) * res_factor_VGA)
    return (float(focal), float(focal), float(f32(f32(cols) / f32(2) - f32(0.5))), float(f32(f32(rows) / f32(2) - f32(0.5))))


def rig_params(n_levels=4, method=PHOTO):
    """Parameters of the rig registration: the class defaults, no sensor-joint mask (that mask lives in alignFrames360)."""
    p = default_params(n_levels=n_levels, method=method, n_sensors_mask=0)
    p.projection = 1
    return p


def _rt_arg(Rt):
    """one 4x4 or eight 4x4 (row-major numpy) -> column-major float32"""
    M = np.asarray(Rt, np.float32)
    if M.ndim == 2:
        return np.ascontiguousarray(M.T).reshape(16)
    return np.ascontiguousarray(M.transpose(0, 2, 1)).reshape(-1, 16)


def inverse4(M):
    A = _rt_arg(M); out = np.zeros(16, np.float32)
    lib().orc_inverse4(_ptr(A), _ptr(out))
    return out.reshape(4, 4).T.copy()


def synth_rig_frame(kind, fid, rows, cols, Rt, cam=None):
    """The 8 pinhole sensor views of the synthetic room from robot frame `fid`; Rt: 8 x 4 x 4 sensor poses in the robot
    frame.  -> (rgb 8 x rows x cols x 3, depth 8 x rows x cols u16 z-depth mm)."""
    cam = cam or rig_camera(rows, cols)
    R = _rt_arg(Rt)
    rgb = np.zeros((8, rows, cols, 3), np.uint8); d = np.zeros((8, rows, cols), np.uint16)
    for s in range(8):
        lib().orc_synth_pinhole_frame_rt(kind, fid, rows, cols, *[np.float32(x) for x in cam], _ptr(R[s]), _ptr(rgb[s]), _ptr(d[s]))
    return rgb, d


def error_robot(src, trg, level, pose, Rt, params, cam):
    """calcPhotoICPError_robot (RPI.h:4905) of one sensor -> (error2, number of terms)."""
    e = C.c_double(); n = C.c_int()
    T = pose_arg(pose); K = _cam(cam); R = _rt_arg(Rt)
    lib().orc_error_robot(src.h, trg.h, level, _ptr(T), _ptr(R), C.byref(params), _ptr(K), C.byref(e), C.byref(n))
    return e.value, n.value


def hessgrad_robot(src, trg, level, pose, Rt, params, cam):
    """calcHessianGradient_robot (RPI.h:5100) of one sensor, PHOTO_CONSISTENCY: H, g float as upstream; Hd, gd double sums."""
    H = np.zeros(36, np.float32); g = np.zeros(6, np.float32)
    Hd = np.zeros(21, np.float64); gd = np.zeros(6, np.float64); cnt = np.zeros(3, np.int32)
    T = pose_arg(pose); K = _cam(cam); R = _rt_arg(Rt)
    lib().orc_hessgrad_robot(src.h, trg.h, level, _ptr(T), _ptr(R), C.byref(params), _ptr(K), _ptr(H), _ptr(g), _ptr(Hd), _ptr(gd), _ptr(cnt))
    return dict(H=H.reshape(6, 6), g=g, Hd=Hd, gd=gd, n_visible=int(cnt[0]), n_photo=int(cnt[1]))


def align_rig(src, trg, Rt, guess, params, cam, faithful=True, accum=ACC_FAITHFUL):
    """RegisterRGBD360::RegisterDensePhotoICP over 8 (source, target) sensor frames.  faithful: new_error is evaluated at
    pose_estim as upstream does (RegisterRGBD360.h:462, 488); False: at the candidate."""
    res = Result()
    sp = (C.c_void_p * 8)(*[f.h for f in src]); tp = (C.c_void_p * 8)(*[f.h for f in trg])
    T = pose_arg(guess); K = _cam(cam); R = _rt_arg(Rt)
    lib().orc_align_rig(sp, tp, _ptr(R), _ptr(T), C.byref(params), _ptr(K), int(bool(faithful)), accum, C.byref(res))
    return res


def synth_gt_pose(kind, src_id, trg_id):
    T = np.zeros(16, np.float64)
    lib().orc_synth_gt_pose(kind, src_id, trg_id, _ptr(T))
    return T.reshape(4, 4).T.copy()


def pinned_vec(fn, a, b=None):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b if b is not None else np.zeros_like(a), np.float32)
    out = np.zeros_like(a)
    lib().orc_pinned_vec(fn, a.size, _ptr(a), _ptr(b), _ptr(out))
    return out


def stitch(sensor_rgb, sensor_depth, Rt_inv, fx=262.5, fy=262.5, cx=159.5, cy=119.5):
    """Frame360::stitchSphericalImage (Frame360.h:386-405, 1099-1148).  sensor_rgb 8 x h x w x 3 u8,
    sensor_depth 8 x h x w u16 mm, Rt_inv 8 x 4 x 4 (row-major numpy matrices)."""
    sensor_rgb = np.ascontiguousarray(sensor_rgb, np.uint8); sensor_depth = np.ascontiguousarray(sensor_depth, np.uint16)
    h, w = sensor_depth.shape[1:]
    cols = 8 * h; rows = int(cols * 0.5 * 60.0 / 180)
    R = np.ascontiguousarray(np.asarray(Rt_inv, np.float32).transpose(0, 2, 1)).reshape(8, 16)    # column-major
    rgb = np.zeros((rows, cols, 3), np.uint8); d = np.zeros((rows, cols), np.uint16)
    lib().orc_stitch(h, w, fx, fy, cx, cy, _ptr(R), _ptr(sensor_rgb), _ptr(sensor_depth), _ptr(rgb), _ptr(d))
    return rgb, d
