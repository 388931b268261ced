"""ctypes loader + thin object wrapper over include/r360.h (one Context per GPU)."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# R360_LIB: development override (kernel variants built side by side); the product is the in-tree library
_SO = os.environ.get("R360_LIB") or os.path.join(_HERE, "librgbd360_b200.so")

R360_MAX_LEVELS = 8
PHOTO_CONSISTENCY, DEPTH_CONSISTENCY, PHOTO_DEPTH = 0, 1, 2
ROLE_SOURCE, ROLE_TARGET, ROLE_BOTH = 1, 2, 3

EXPORTS = [
    "r360_default_params", "r360_last_error", "r360_create", "r360_destroy", "r360_set_frames",
    "r360_set_frames_dev", "r360_set_frames_f32", "r360_register_pairs", "r360_eval_error",
    "r360_eval_hessgrad", "r360_dump_level", "r360_dump_source_level", "r360_dump_warp",
    "r360_synth_frames_dev", "r360_synth_frames", "r360_synth_gt_pose", "r360_device_alloc",
    "r360_device_free", "r360_synchronize", "r360_last_device_ms", "r360_kernel_launches",
    "r360_last_pass_stats", "r360_version", "r360_index_stats", "r360_register_host_pairs",
    "r360_default_rig", "r360_frame360_parse", "r360_stitch_frames", "r360_eval_error_occ",
    "r360_default_params_pinhole", "r360_set_camera", "r360_eval_error_pinhole",
    "r360_allgather_results", "r360_host_alloc", "r360_host_free", "r360_register_rig_pairs", "r360_eval_rig",
]


class R360Error(RuntimeError):
    pass


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


class Rig(C.Structure):
    """Calib360 (Calib360.h:69-131): shared pinhole intrinsics + inverse extrinsics of the 8 sensors."""
    _fields_ = [("sensor_rows", C.c_int32), ("sensor_cols", C.c_int32), ("fx", C.c_float), ("fy", C.c_float),
                ("cx", C.c_float), ("cy", C.c_float), ("Rt_inv", (C.c_float * 16) * 8)]


RESULT_DTYPE = np.dtype([
    ("pose", np.float32, 16), ("hessian", np.float32, 36), ("gradient", np.float32, 6),
    ("sso", np.float32), ("n_visible", np.int32), ("final_error", np.float64),
    ("final_err2", np.float64), ("final_n_valid", np.int32), ("status", np.int32),
    ("iters", np.int32, R360_MAX_LEVELS), ("passes", np.int32, R360_MAX_LEVELS),
    ("pair_id", np.int32), ("reserved", np.int32)])
assert RESULT_DTYPE.itemsize == C.sizeof(Result)


def build_native(verbose=False):
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc")], capture_output=True, text=True)
    if verbose or out.returncode:
        print(out.stdout, out.stderr)
    if out.returncode:
        raise R360Error("building librgbd360_b200.so failed")
    return _SO


_lib = None


def lib():
    """Load librgbd360_b200.so; fails loudly when the CUDA extension is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise R360Error(f"{_SO} not found: build it with __graft_entry__.build() "
                        "(there is no CPU fallback for this path)")
    L = C.CDLL(_SO)
    vp, i32, f32 = C.c_void_p, C.c_int, C.c_float
    L.r360_default_params.argtypes = [C.POINTER(Params)]
    L.r360_last_error.restype = C.c_char_p
    L.r360_last_error.argtypes = [vp]
    L.r360_create.argtypes = [C.POINTER(vp), i32, i32, i32, i32, i32, C.POINTER(Params)]
    L.r360_destroy.argtypes = [vp]
    L.r360_destroy.restype = None
    L.r360_set_frames.argtypes = [vp, i32, i32, vp, vp, vp]
    L.r360_set_frames_dev.argtypes = [vp, i32, i32, vp, vp, vp]
    L.r360_set_frames_f32.argtypes = [vp, i32, i32, vp, vp, vp]
    L.r360_register_pairs.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    L.r360_register_host_pairs.argtypes = [vp, i32, vp, vp, vp, vp]
    L.r360_eval_error.argtypes = [vp, i32, i32, i32, vp, C.POINTER(C.c_double), C.POINTER(C.c_int32)]
    L.r360_eval_error_occ.argtypes = [vp, i32, i32, i32, vp, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                      C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_double)]
    L.r360_eval_error_pinhole.argtypes = L.r360_eval_error_occ.argtypes
    L.r360_default_params_pinhole.argtypes = [C.POINTER(Params)]
    L.r360_set_camera.argtypes = [vp, f32, f32, f32, f32]
    L.r360_eval_hessgrad.argtypes = [vp, i32, i32, i32, vp, vp, vp, C.POINTER(C.c_int32)]
    L.r360_dump_level.argtypes = [vp, i32, i32] + [vp] * 6
    L.r360_dump_source_level.argtypes = [vp, i32, i32, vp, vp]
    L.r360_dump_warp.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, vp]
    L.r360_synth_frames_dev.argtypes = [vp, i32, i32, i32, vp, vp]
    L.r360_synth_frames.argtypes = [vp, i32, i32, i32, vp, vp]
    L.r360_synth_gt_pose.argtypes = [i32, i32, i32, vp]
    L.r360_synth_gt_pose.restype = None
    L.r360_index_stats.argtypes = [vp, i32, i32, i32, vp, vp]
    L.r360_device_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    L.r360_device_free.argtypes = [vp, vp]
    L.r360_synchronize.argtypes = [vp]
    L.r360_last_device_ms.argtypes = [vp]
    L.r360_last_device_ms.restype = f32
    L.r360_kernel_launches.argtypes = [vp]
    L.r360_kernel_launches.restype = C.c_int64
    L.r360_last_pass_stats.argtypes = [vp, C.POINTER(f32), C.POINTER(C.c_int32), C.POINTER(C.c_double)]
    L.r360_default_rig.argtypes = [C.POINTER(Rig)]
    L.r360_default_rig.restype = None
    L.r360_frame360_parse.argtypes = [vp, C.c_size_t, C.POINTER(C.c_int32), C.POINTER(C.c_int32), vp, C.c_size_t, vp, C.c_size_t]
    L.r360_stitch_frames.argtypes = [vp, C.POINTER(Rig), i32, i32, vp, vp, vp, vp, vp]
    L.r360_allgather_results.argtypes = [vp, vp, vp, i32, i32, vp]
    L.r360_register_rig_pairs.argtypes = [vp, i32, vp, vp, vp, vp, i32, vp]
    L.r360_eval_rig.argtypes = [vp, i32, i32, i32, vp, vp, C.POINTER(C.c_double), vp, vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.r360_host_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp), C.POINTER(C.c_int32)]
    L.r360_host_free.argtypes = [vp, vp]
    _lib = L
    return L


def make_rig(Rt_inv=None, sensor_rows=240, sensor_cols=320, fx=262.5, fy=262.5, cx=159.5, cy=119.5):
    """r360_rig from 8 inverse extrinsic matrices (4x4, row-major numpy) -- Calib360::Rt_inv."""
    rig = Rig()
    lib().r360_default_rig(C.byref(rig))
    rig.sensor_rows, rig.sensor_cols = sensor_rows, sensor_cols
    rig.fx, rig.fy, rig.cx, rig.cy = fx, fy, cx, cy
    if Rt_inv is not None:
        M = np.asarray(Rt_inv, np.float32).reshape(8, 4, 4)
        for s in range(8):
            col = np.ascontiguousarray(M[s].T).reshape(16)
            for k in range(16):
                rig.Rt_inv[s][k] = float(col[k])
    return rig


def sphere_shape(sensor_rows):
    """(rows, cols) of the stitched sphere image (Frame360.h:391-392)."""
    cols = 8 * sensor_rows
    return int(cols * 0.5 * 60.0 / 180), cols


def frame360_parse(data):
    """Frame360::loadFrame (Frame360.h:231-266) on the bytes of a .bin: -> (rgb 8xhxwx3 u8, depth 8xhxw u16)."""
    buf = np.frombuffer(data, np.uint8)
    r, c = C.c_int32(), C.c_int32()
    rc = lib().r360_frame360_parse(buf.ctypes.data_as(C.c_void_p), buf.size, C.byref(r), C.byref(c), None, 0, None, 0)
    if rc:
        raise R360Error(f"r360_frame360_parse: not a Frame360 archive (error {rc})")
    rgb = np.zeros((8, r.value, c.value, 3), np.uint8); dep = np.zeros((8, r.value, c.value), np.uint16)
    rc = lib().r360_frame360_parse(buf.ctypes.data_as(C.c_void_p), buf.size, C.byref(r), C.byref(c),
                                   rgb.ctypes.data_as(C.c_void_p), rgb.size, dep.ctypes.data_as(C.c_void_p), dep.size)
    if rc:
        raise R360Error(f"r360_frame360_parse failed (error {rc})")
    return rgb, dep


def default_params(**kw):
    p = Params()
    lib().r360_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def pinhole_params(**kw):
    """r360_default_params_pinhole: the constants of the pinhole alignFrames (RPI.h:4304-4309)."""
    p = Params()
    lib().r360_default_params_pinhole(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def pose_to_colmajor(pose):
    """4x4 matrix -> 16 floats column-major (Eigen::Matrix4f layout)."""
    return np.ascontiguousarray(np.asarray(pose, np.float32).reshape(4, 4).T).reshape(16)


def pose_from_colmajor(buf):
    return np.array(buf, np.float32).reshape(4, 4).T.copy()


def synth_gt_pose(kind, src_id, trg_id):
    T = np.zeros(16, np.float64)
    lib().r360_synth_gt_pose(kind, src_id, trg_id, T.ctypes.data_as(C.c_void_p))
    return T.reshape(4, 4).T.copy()


def _p(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """One r360_ctx: frame slots + pair batch on one GPU."""

    def __init__(self, rows, cols, max_frames, max_pairs, params=None, device=0):
        self.L = lib()
        self.params = params if params is not None else default_params()
        self.rows, self.cols = rows, cols
        self.max_frames, self.max_pairs = max_frames, max_pairs
        self.h = C.c_void_p()
        rc = self.L.r360_create(C.byref(self.h), device, rows, cols, max_frames, max_pairs, C.byref(self.params))
        if rc:
            raise R360Error(f"r360_create failed ({rc}): {self.L.r360_last_error(None).decode()}")

    def close(self):
        if getattr(self, "h", None) and self.h:
            self.L.r360_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise R360Error(f"r360 error {rc}: {self.L.r360_last_error(self.h).decode()}")

    # ---- frames
    def _check_frames(self, first, n, roles):
        if n < 0 or first < 0 or first + n > self.max_frames:
            raise ValueError(f"frames [{first}, {first + n}) outside the {self.max_frames} slots of the context")
        if roles is not None and roles.size != n:
            raise ValueError(f"roles has {roles.size} entries for {n} frames")

    def set_frames(self, first, rgb, depth, roles=None):
        rgb = np.ascontiguousarray(rgb, np.uint8)
        depth = np.asarray(depth)
        n = rgb.shape[0] if rgb.ndim == 4 else 1
        r = None if roles is None else np.ascontiguousarray(roles, np.uint8)
        # the C library reads n * rows * cols [* 3] elements from raw pointers: check here, raise here
        if rgb.shape[-3:] != (self.rows, self.cols, 3) or rgb.ndim not in (3, 4):
            raise ValueError(f"rgb has shape {rgb.shape}, expected [n x] {self.rows} x {self.cols} x 3")
        if depth.size != n * self.rows * self.cols or depth.shape[-2:] != (self.rows, self.cols):
            raise ValueError(f"depth has shape {depth.shape}, expected [{n} x] {self.rows} x {self.cols}")
        self._check_frames(first, n, r)
        if depth.dtype == np.uint16:
            d = np.ascontiguousarray(depth)
            self._ck(self.L.r360_set_frames(self.h, first, n, _p(rgb), _p(d), _p(r)))
        else:
            d = np.ascontiguousarray(depth, np.float32)
            self._ck(self.L.r360_set_frames_f32(self.h, first, n, _p(rgb), _p(d), _p(r)))

    def set_frames_ptr(self, first, n, rgb_ptr, depth_ptr, roles=None, device=False):
        """Raw-pointer form (host pinned buffers or device pointers)."""
        r = None if roles is None else np.ascontiguousarray(roles, np.uint8)
        self._check_frames(first, n, r)
        fn = self.L.r360_set_frames_dev if device else self.L.r360_set_frames
        self._ck(fn(self.h, first, n, _p(rgb_ptr), _p(depth_ptr), _p(r)))

    def stitch_frames(self, rig, first, sensor_rgb, sensor_depth, roles=None, want_sphere=True):
        """Frame360::stitchSphericalImage on the device + set*Frame of the stitched spheres.
        sensor_rgb: n x 8 x h x w x 3 u8, sensor_depth: n x 8 x h x w u16 (mm)."""
        sensor_rgb = np.ascontiguousarray(sensor_rgb, np.uint8); sensor_depth = np.ascontiguousarray(sensor_depth, np.uint16)
        n = sensor_rgb.shape[0]
        r = None if roles is None else np.ascontiguousarray(roles, np.uint8)
        if sensor_rgb.shape != (n, 8, rig.sensor_rows, rig.sensor_cols, 3) or sensor_depth.shape != (n, 8, rig.sensor_rows, rig.sensor_cols):
            raise ValueError(f"sensor images have shapes {sensor_rgb.shape} / {sensor_depth.shape}, expected "
                             f"n x 8 x {rig.sensor_rows} x {rig.sensor_cols} [x 3]")
        self._check_frames(first, n, r)
        srgb = np.zeros((n, self.rows, self.cols, 3), np.uint8) if want_sphere else None
        sdep = np.zeros((n, self.rows, self.cols), np.uint16) if want_sphere else None
        self._ck(self.L.r360_stitch_frames(self.h, C.byref(rig), first, n, _p(sensor_rgb), _p(sensor_depth), _p(r),
                                           _p(srgb), _p(sdep)))
        return srgb, sdep

    def synth_frames(self, kind, first_id, n):
        rgb = np.zeros((n, self.rows, self.cols, 3), np.uint8)
        d = np.zeros((n, self.rows, self.cols), np.uint16)
        self._ck(self.L.r360_synth_frames(self.h, kind, first_id, n, _p(rgb), _p(d)))
        return rgb, d

    def synth_frames_dev(self, kind, first_id, n, rgb_ptr, depth_ptr):
        self._ck(self.L.r360_synth_frames_dev(self.h, kind, first_id, n, _p(rgb_ptr), _p(depth_ptr)))

    # ---- registration
    @staticmethod
    def _check_out(out, n):
        if out is None:
            return np.zeros(n, RESULT_DTYPE)
        if not isinstance(out, np.ndarray) or out.dtype != RESULT_DTYPE or out.size < n or not out.flags.c_contiguous:
            raise ValueError(f"out must be a contiguous array of >= {n} RESULT_DTYPE records")
        return out

    def register_pairs(self, src_idx, trg_idx, init_pose=None, trace=False, out=None):
        s = np.ascontiguousarray(src_idx, np.int32)
        t = np.ascontiguousarray(trg_idx, np.int32)
        n = s.size
        if t.size != n:
            raise ValueError(f"src_idx has {n} entries, trg_idx {t.size}")
        if n > self.max_pairs:
            raise ValueError(f"{n} pairs exceed max_pairs = {self.max_pairs}")
        res = self._check_out(out, n)
        ip = None
        if init_pose is not None:
            ip = np.ascontiguousarray(init_pose, np.float32)
            if ip.size != 16 * n:
                raise ValueError(f"init_pose has {ip.size} floats, expected {n} x 16")
            ip = ip.reshape(n, 16)
        tr = None
        if trace:
            per = (2 * self.params.max_iters + 2) if self.params.projection == 1 else (self.params.max_iters + 2)
            tr = (IterRecord * (n * self.params.n_levels * per))()
        self._ck(self.L.r360_register_pairs(self.h, n, _p(s), _p(t), _p(ip), _p(res),
                                            C.cast(tr, C.c_void_p) if trace else None))
        return (res, tr) if trace else res

    def register_host_pairs(self, rgb, depth_mm, n_pairs=None, init_pose=None, out=None):
        """setTargetFrame + setSourceFrame + alignFrames360 of host frames in one pipelined call:
        frame 2p = target, frame 2p+1 = source of pair p.  rgb / depth_mm: arrays or raw host pointers."""
        if n_pairs is None:
            n_pairs = rgb.shape[0] // 2
        if not isinstance(rgb, int):
            rgb = np.ascontiguousarray(rgb, np.uint8)
            depth_mm = np.ascontiguousarray(depth_mm, np.uint16)
            if rgb.shape != (2 * n_pairs, self.rows, self.cols, 3) or depth_mm.shape != (2 * n_pairs, self.rows, self.cols):
                raise ValueError(f"frames have shapes {rgb.shape} / {depth_mm.shape}, expected "
                                 f"{2 * n_pairs} x {self.rows} x {self.cols} [x 3]")
        if n_pairs > self.max_pairs or 2 * n_pairs > self.max_frames:
            raise ValueError(f"{n_pairs} pairs need max_pairs >= {n_pairs} and max_frames >= {2 * n_pairs}")
        res = self._check_out(out, n_pairs)
        ip = None
        if init_pose is not None:
            ip = np.ascontiguousarray(init_pose, np.float32)
            if ip.size != 16 * n_pairs:
                raise ValueError(f"init_pose has {ip.size} floats, expected {n_pairs} x 16")
            ip = ip.reshape(n_pairs, 16)
        self._ck(self.L.r360_register_host_pairs(self.h, n_pairs, _p(rgb), _p(depth_mm), _p(ip), _p(res)))
        return res

    def eval_error(self, src, trg, level, pose):
        e2, n = C.c_double(), C.c_int32()
        T = pose_to_colmajor(pose)
        self._ck(self.L.r360_eval_error(self.h, src, trg, level, _p(T), C.byref(e2), C.byref(n)))
        return e2.value, n.value

    def eval_error_occ(self, src, trg, level, pose):
        """errorPhotoICP_sphereOcc1 / Occ2 (ctx created with occlusion 1 / 2) -> dict(photo, depth, n_photo, n_depth, error)."""
        pr, dr, e = C.c_double(), C.c_double(), C.c_double()
        npv, ndv = C.c_int32(), C.c_int32()
        T = pose_to_colmajor(pose)
        self._ck(self.L.r360_eval_error_occ(self.h, src, trg, level, _p(T), C.byref(pr), C.byref(dr), C.byref(npv),
                                            C.byref(ndv), C.byref(e)))
        return dict(photo=pr.value, depth=dr.value, n_photo=npv.value, n_depth=ndv.value, error=e.value)

    def set_camera(self, fx, fy, ox, oy):
        """setCameraMatrix (RPI.h:254) of a pinhole context."""
        self._ck(self.L.r360_set_camera(self.h, fx, fy, ox, oy))

    def eval_error_pinhole(self, src, trg, level, pose):
        """errorPhotoICP (RPI.h:560) -> dict(photo, depth, n_photo, n_depth, error)."""
        pr, dr, e = C.c_double(), C.c_double(), C.c_double()
        npv, ndv = C.c_int32(), C.c_int32()
        T = pose_to_colmajor(pose)
        self._ck(self.L.r360_eval_error_pinhole(self.h, src, trg, level, _p(T), C.byref(pr), C.byref(dr), C.byref(npv),
                                                C.byref(ndv), C.byref(e)))
        return dict(photo=pr.value, depth=dr.value, n_photo=npv.value, n_depth=ndv.value, error=e.value)

    def eval_hessgrad(self, src, trg, level, pose):
        H = np.zeros(36, np.float32); g = np.zeros(6, np.float32); nv = C.c_int32()
        T = pose_to_colmajor(pose)
        self._ck(self.L.r360_eval_hessgrad(self.h, src, trg, level, _p(T), _p(H), _p(g), C.byref(nv)))
        return H.reshape(6, 6), g, nv.value

    def dump_level(self, frame, level, grads=True):
        r, c = self.rows >> level, self.cols >> level
        names = ["gray", "depth"] + (["ggx", "ggy", "dgx", "dgy"] if grads else [])
        out = {k: np.zeros((r, c), np.float32) for k in names}
        args = [_p(out[k]) if k in out else None for k in ["gray", "depth", "ggx", "ggy", "dgx", "dgy"]]
        self._ck(self.L.r360_dump_level(self.h, frame, level, *args))
        return out

    def dump_source_level(self, frame, level):
        r, c = self.rows >> level, self.cols >> level
        g = np.zeros((r, c), np.float32); d = np.zeros((r, c), np.float32)
        self._ck(self.L.r360_dump_source_level(self.h, frame, level, _p(g), _p(d)))
        return dict(gray=g, depth=d)

    def dump_warp(self, src, trg, level, pose):
        n = (self.rows >> level) * (self.cols >> level)
        ri = np.zeros(n, np.int32); ci = np.zeros(n, np.int32)
        vp = np.zeros(n, np.uint8); vd = np.zeros(n, np.uint8)
        T = pose_to_colmajor(pose)
        self._ck(self.L.r360_dump_warp(self.h, src, trg, level, _p(T), _p(ri), _p(ci), _p(vp), _p(vd)))
        return ri, ci, vp, vd

    def index_stats(self, src, trg, level, pose):
        """Packed vs scalar pinned index path (see include/r360.h): dict(valid, scalar, mismatch)."""
        out = np.zeros(3, np.uint64)
        T = pose_to_colmajor(pose)
        self._ck(self.L.r360_index_stats(self.h, src, trg, level, _p(T), _p(out)))
        return dict(valid=int(out[0]), scalar=int(out[1]), mismatch=int(out[2]))

    # ---- the 8-sensor rig (RegisterRGBD360::RegisterDensePhotoICP)
    @staticmethod
    def _rt8(Rt):
        M = np.asarray(Rt, np.float32)
        if M.shape != (8, 4, 4):
            raise ValueError(f"Rt has shape {M.shape}, expected 8 x 4 x 4 (calib->Rt_)")
        return np.ascontiguousarray(M.transpose(0, 2, 1)).reshape(8, 16)              # column-major each

    def register_rig_pairs(self, src_first, trg_first, Rt, init_pose=None, faithful=True, out=None):
        """r360_register_rig_pairs: rig frames occupy 8 consecutive slots; src_first / trg_first = slot of sensor 0 of
        frame2 / frame1.  Rt: 8 x 4 x 4 sensor poses (row-major numpy)."""
        s = np.ascontiguousarray(src_first, np.int32); t = np.ascontiguousarray(trg_first, np.int32)
        n = s.size
        if t.size != n or n > self.max_pairs:
            raise ValueError(f"{n} / {t.size} rig pairs (max_pairs = {self.max_pairs})")
        res = self._check_out(out, n)
        ip = None
        if init_pose is not None:
            ip = np.ascontiguousarray(init_pose, np.float32)
            if ip.size != 16 * n:
                raise ValueError(f"init_pose has {ip.size} floats, expected {n} x 16")
        R = self._rt8(Rt)
        self._ck(self.L.r360_register_rig_pairs(self.h, n, _p(s), _p(t), _p(R), _p(ip), int(bool(faithful)), _p(res)))
        return res

    def eval_rig(self, src_first, trg_first, level, pose, Rt):
        """r360_eval_rig -> dict(error2, H 6x6, g, n_visible, n_terms): the rig's summed *_robot functions at `pose`."""
        e = C.c_double(); nv = C.c_int32(); nt = C.c_int32()
        H = np.zeros(36, np.float32); g = np.zeros(6, np.float32)
        T = pose_to_colmajor(pose); R = self._rt8(Rt)
        self._ck(self.L.r360_eval_rig(self.h, src_first, trg_first, level, _p(T), _p(R), C.byref(e), _p(H), _p(g), C.byref(nv), C.byref(nt)))
        return dict(error2=e.value, H=H.reshape(6, 6), g=g, n_visible=nv.value, n_terms=nt.value)

    # ---- multi-GPU exchange and host staging memory
    def allgather_results(self, nccl_comm, local, n_ranks):
        """r360_allgather_results: `local` (RESULT_DTYPE records, the same count on every rank) -> all ranks' records,
        rank-major, on the host.  nccl_comm: the caller's ncclComm_t for this ctx's device (an int / c_void_p)."""
        local = np.ascontiguousarray(local)
        if local.dtype != RESULT_DTYPE:
            raise ValueError("local must be an array of RESULT_DTYPE records")
        out = np.zeros(n_ranks * local.size, RESULT_DTYPE)
        comm = nccl_comm if isinstance(nccl_comm, C.c_void_p) else C.c_void_p(int(nccl_comm))
        self._ck(self.L.r360_allgather_results(self.h, comm, _p(local), local.size, n_ranks, _p(out)))
        return out

    def host_alloc(self, nbytes):
        """r360_host_alloc: pinned host memory on the GPU's NUMA node -> (address, node or -1, note)."""
        p, node = C.c_void_p(), C.c_int32(-1)
        self._ck(self.L.r360_host_alloc(self.h, nbytes, C.byref(p), C.byref(node)))
        return int(p.value), int(node.value), self.L.r360_last_error(self.h).decode()

    def host_free(self, addr):
        self._ck(self.L.r360_host_free(self.h, C.c_void_p(addr)))

    # ---- plumbing
    def synchronize(self):
        self._ck(self.L.r360_synchronize(self.h))

    def last_device_ms(self):
        return float(self.L.r360_last_device_ms(self.h))

    def kernel_launches(self):
        return int(self.L.r360_kernel_launches(self.h))

    def last_pass_stats(self):
        ms, n, b = C.c_float(), C.c_int32(), C.c_double()
        self._ck(self.L.r360_last_pass_stats(self.h, C.byref(ms), C.byref(n), C.byref(b)))
        return dict(ms=ms.value, launches=n.value, alg_bytes=b.value)
