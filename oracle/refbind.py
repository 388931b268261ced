"""ctypes binding of oracle/_ref/librpi_ref.so -- the reference's own RegisterPhotoICP.h compiled
against the from-scratch third-party stand-ins of oracle/refshim/ (see oracle/ref_harness.cpp).

TEST INFRASTRUCTURE ONLY.  The library can only be (re)built where /root/reference exists; the
prebuilt .so travels to the GPU box (oracle/_ref/ is git-ignored, not gpurun-ignored).
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = {False: os.path.join(_HERE, "_ref", "librpi_ref.so"),          # glibc asinf/atan2f/sinf/cosf
       True: os.path.join(_HERE, "_ref", "librpi_ref_pinned.so")}    # sphere_math.h sequences (what the GPU runs)
REFERENCE_DIR = "/root/reference"


def available():
    return all(os.path.exists(p) for p in _SO.values()) or os.path.isdir(os.path.join(REFERENCE_DIR, "include"))


def build():
    if os.path.isdir(os.path.join(REFERENCE_DIR, "include")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
    return _SO


_libs = {}


def lib(pinned=False):
    pinned = bool(pinned)
    if pinned not in _libs:
        if not os.path.exists(_SO[pinned]):
            build()
        L = C.CDLL(_SO[pinned])
        assert L.ref_pinned_math() == int(pinned)
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float, C.c_float]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_set_source.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.ref_set_target.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.ref_level.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 6
        L.ref_align.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 7 + [C.c_int]
        L.ref_error.restype = C.c_double
        L.ref_error.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.ref_hessgrad.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
        L.ref_error_occ.restype = C.c_double
        L.ref_error_occ.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.ref_hessgrad_occ.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
        L.ref_set_camera.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float]
        L.ref_align_pinhole.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 4
        L.ref_error_pinhole.restype = C.c_double
        L.ref_error_pinhole.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.ref_hessgrad_pinhole.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_lut.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ref_error_robot.restype = C.c_double
        L.ref_error_robot.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_hessgrad_robot.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_rig_align.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _libs[pinned] = L
    return _libs[pinned]


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _pose_arg(pose):
    if pose is None:
        pose = np.eye(4)
    return np.ascontiguousarray(np.asarray(pose, dtype=np.float32).reshape(4, 4).T).reshape(16)


class Reference:
    """One RegisterPhotoICP instance of the reference (RPI.h:85)."""

    def __init__(self, n_levels=4, min_depth=0.3, max_depth=6.0, std_photo=6.0 / 255, std_depth=0.2, pinned=False):
        self.n_levels = n_levels
        self.L = lib(pinned)
        self.h = self.L.ref_create(n_levels, min_depth, max_depth, np.float32(std_photo), std_depth)
        self.rows = self.cols = 0

    def close(self):
        if getattr(self, "h", None):
            self.L.ref_destroy(self.h)
            self.h = None

    __del__ = close

    def set_source(self, rgb, depth_mm):
        rgb = np.ascontiguousarray(rgb, np.uint8); d = np.ascontiguousarray(depth_mm, np.uint16)
        self.rows, self.cols = d.shape
        self.L.ref_set_source(self.h, _ptr(rgb), _ptr(d), self.rows, self.cols)

    def set_target(self, rgb, depth_mm):
        rgb = np.ascontiguousarray(rgb, np.uint8); d = np.ascontiguousarray(depth_mm, np.uint16)
        self.rows, self.cols = d.shape
        self.L.ref_set_target(self.h, _ptr(rgb), _ptr(d), self.rows, self.cols)

    def level(self, which, level):
        r, c = self.rows >> level, self.cols >> level
        names = ["gray", "depth"] + (["ggx", "ggy", "dgx", "dgy"] if which == 1 else [])
        out = {n: np.zeros((r, c), np.float32) for n in names}
        args = [_ptr(out[n]) if n in out else None for n in ["gray", "depth", "ggx", "ggy", "dgx", "dgy"]]
        assert self.L.ref_level(self.h, which, level, *args) == 0
        return out

    def align(self, guess=None, method=2, occlusion=0):
        T = _pose_arg(guess)
        pose = np.zeros(16, np.float32); H = np.zeros(36, np.float32); g = np.zeros(6, np.float32)
        sso = np.zeros(1, np.float32); iters = np.zeros(self.n_levels, np.int32)
        cap = self.n_levels * 24
        e2 = np.zeros(cap, np.float64); nv = np.zeros(cap, np.int32)
        n = self.L.ref_align(self.h, _ptr(T), method, occlusion, _ptr(pose), _ptr(H), _ptr(g), _ptr(sso), _ptr(iters),
                            _ptr(e2), _ptr(nv), cap)
        ill = n < 0
        if ill:
            n = -1 - n
        return dict(pose=pose.reshape(4, 4).T.copy(), H=H.reshape(6, 6).T.copy(), g=g, sso=float(sso[0]),
                    iters=iters, err2=e2[:n].copy(), n_valid=nv[:n].copy(), ill_posed=ill)

    def error(self, level, pose, method=2):
        e2, n = C.c_double(), C.c_int()
        e = self.L.ref_error(self.h, level, _ptr(_pose_arg(pose)), method, C.byref(e2), C.byref(n))
        return e, e2.value, n.value

    def hessgrad(self, level, pose, method=2):
        H = np.zeros(36, np.float32); g = np.zeros(6, np.float32); sso = C.c_float()
        self.L.ref_hessgrad(self.h, level, _ptr(_pose_arg(pose)), method, _ptr(H), _ptr(g), C.byref(sso))
        return H.reshape(6, 6).T.copy(), g, sso.value

    def error_occ(self, level, pose, method=2, occlusion=1):
        """errorPhotoICP_sphereOcc1/2 -> (return value, avPhotoResidual, avDepthResidual)."""
        av = np.zeros(2, np.float64)
        e = self.L.ref_error_occ(self.h, level, _ptr(_pose_arg(pose)), method, occlusion, _ptr(av))
        return e, float(av[0]), float(av[1])

    def hessgrad_occ(self, level, pose, method=2, occlusion=1):
        H = np.zeros(36, np.float32); g = np.zeros(6, np.float32); sso = C.c_float()
        self.L.ref_hessgrad_occ(self.h, level, _ptr(_pose_arg(pose)), method, occlusion, _ptr(H), _ptr(g), C.byref(sso))
        return H.reshape(6, 6).T.copy(), g, sso.value

    # ---- pinhole path (RPI.h:254, 560, 776, 4254)
    def set_camera(self, fx, fy, ox, oy):
        self.L.ref_set_camera(self.h, fx, fy, ox, oy)

    def align_pinhole(self, guess=None, method=2):
        T = _pose_arg(guess)
        pose = np.zeros(16, np.float32); H = np.zeros(36, np.float32); g = np.zeros(6, np.float32)
        iters = np.zeros(self.n_levels, np.int32)
        ill = self.L.ref_align_pinhole(self.h, _ptr(T), method, _ptr(pose), _ptr(H), _ptr(g), _ptr(iters))
        return dict(pose=pose.reshape(4, 4).T.copy(), H=H.reshape(6, 6).T.copy(), g=g, iters=iters, ill_posed=bool(ill))

    def error_pinhole(self, level, pose, method=2):
        av = np.zeros(2, np.float64)
        e = self.L.ref_error_pinhole(self.h, level, _ptr(_pose_arg(pose)), method, _ptr(av))
        return e, float(av[0]), float(av[1])

    def hessgrad_pinhole(self, level, pose, method=2):
        H = np.zeros(36, np.float32); g = np.zeros(6, np.float32)
        self.L.ref_hessgrad_pinhole(self.h, level, _ptr(_pose_arg(pose)), method, _ptr(H), _ptr(g))
        return H.reshape(6, 6).T.copy(), g

    # ---- the 8-sensor rig: calcPhotoICPError_robot / calcHessianGradient_robot of this sensor (RPI.h:4905, 5100)
    def error_robot(self, level, pose, Rt, method=0):
        return self.L.ref_error_robot(self.h, level, _ptr(_pose_arg(pose)), _ptr(_pose_arg(Rt)), method)

    def hessgrad_robot(self, level, pose, Rt, method=0):
        H = np.zeros(36, np.float32); g = np.zeros(6, np.float32)
        self.L.ref_hessgrad_robot(self.h, level, _ptr(_pose_arg(pose)), _ptr(_pose_arg(Rt)), method, _ptr(H), _ptr(g))
        return H.reshape(6, 6).T.copy(), g

    def lut(self):
        n = self.L.ref_lut(self.h, None, 0)
        out = np.zeros((n, 3), np.float32)
        self.L.ref_lut(self.h, _ptr(out), n)
        return out


# ---------------------------------------------------------------- ingest: the reference's own stitch
_SO_STITCH = {False: os.path.join(_HERE, "_ref", "librpi_ref_stitch.so"),
              True: os.path.join(_HERE, "_ref", "librpi_ref_stitch_pinned.so")}
_stitch_libs = {}


def stitch_available():
    return all(os.path.exists(p) for p in _SO_STITCH.values()) or os.path.isdir(os.path.join(REFERENCE_DIR, "include"))


def stitch_lib(pinned=False):
    pinned = bool(pinned)
    if pinned not in _stitch_libs:
        if not os.path.exists(_SO_STITCH[pinned]):
            build()
        L = C.CDLL(_SO_STITCH[pinned])
        L.refstitch_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_char_p] + [C.c_void_p] * 4 + [C.POINTER(C.c_int)] * 2
        L.refstitch_camera.argtypes = [C.c_void_p]
        _stitch_libs[pinned] = L
    return _stitch_libs[pinned]


def stitch(sensor_rgb, sensor_depth, Rt_inv=None, extrinsics_dir=None, pinned=False):
    """Frame360::stitchSphericalImage + stitchImage as the reference wrote them (oracle/ref_stitch_harness.cpp), with
    Calib360's camera matrix.  sensor_rgb 8 x h x w x 3 u8, sensor_depth 8 x h x w u16.  Rt_inv: 8 x 4 x 4 (row-major
    numpy) put straight into Calib360::Rt_inv, or None: Calib360::loadExtrinsicCalibration(extrinsics_dir or the
    reference's own Calibration/Extrinsics).  -> (sphere rgb, sphere depth u16, the Rt_inv used as 8 x 4 x 4)."""
    L = stitch_lib(pinned)
    sensor_rgb = np.ascontiguousarray(sensor_rgb, np.uint8); sensor_depth = np.ascontiguousarray(sensor_depth, np.uint16)
    h, w = sensor_depth.shape[1:]
    cols = 8 * h
    rows = int(cols * 0.5 * 60.0 / 180)
    rgb = np.zeros((rows, cols, 3), np.uint8); d = np.zeros((rows, cols), np.uint16)
    used = np.zeros((8, 16), np.float32)
    rin = None
    if Rt_inv is not None:
        rin = np.ascontiguousarray(np.asarray(Rt_inv, np.float32).reshape(8, 4, 4).transpose(0, 2, 1)).reshape(8, 16)
    r, c = C.c_int(), C.c_int()
    L.refstitch_run(sensor_rgb.ctypes.data, sensor_depth.ctypes.data, h, w,
                    None if extrinsics_dir is None else extrinsics_dir.encode(),
                    None if rin is None else rin.ctypes.data, rgb.ctypes.data, d.ctypes.data, used.ctypes.data,
                    C.byref(r), C.byref(c))
    assert (r.value, c.value) == (rows, cols)
    return rgb, d, used.reshape(8, 4, 4).transpose(0, 2, 1).copy()


def stitch_camera(pinned=False):
    out = np.zeros(4, np.float32)
    stitch_lib(pinned).refstitch_camera(out.ctypes.data)
    return tuple(float(x) for x in out)


def rig_align(rgb1, d1, rgb2, d2, Rt, guess=None, method=0, pinned=False):
    """RegisterRGBD360::RegisterDensePhotoICP as the reference wrote it (cut out verbatim at build time, see
    oracle/ref_harness.cpp): frame1 = target (8 x h x w [x 3]), frame2 = source, Rt 8 x 4 x 4 row-major numpy.
    -> dict(ok, pose 4x4, info 6x6).  Its OpenMP reduction makes the accept decisions depend on thread arrival order."""
    L = lib(pinned)
    rgb1 = np.ascontiguousarray(rgb1, np.uint8); rgb2 = np.ascontiguousarray(rgb2, np.uint8)
    d1 = np.ascontiguousarray(d1, np.uint16); d2 = np.ascontiguousarray(d2, np.uint16)
    h, w = d1.shape[1:]
    R = np.ascontiguousarray(np.asarray(Rt, np.float32).reshape(8, 4, 4).transpose(0, 2, 1)).reshape(8, 16)
    pose = np.zeros(16, np.float32); info = np.zeros(36, np.float32)
    ok = L.ref_rig_align(_ptr(rgb1), _ptr(d1), _ptr(rgb2), _ptr(d2), h, w, _ptr(R), _ptr(_pose_arg(guess)), method, _ptr(pose), _ptr(info))
    return dict(ok=bool(ok), pose=pose.reshape(4, 4).T.copy(), info=info.reshape(6, 6).T.copy())
