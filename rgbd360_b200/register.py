"""Python mirror of the reference class surface (RegisterPhotoICP, RPI.h:201-288, 480-516, 4519).

Same method names and call-order contract as the C++ class (setNumPyr before set*Frame;
setTargetFrame / setSourceFrame in either order; alignFrames360; getters), one pair per
instance, backed by a two-slot Context.  For throughput use Context.register_pairs directly.
"""
import numpy as np
from .native import (Context, default_params, PHOTO_CONSISTENCY, ROLE_SOURCE, ROLE_TARGET,
                     pose_from_colmajor, pose_to_colmajor)


class RegisterPhotoICP:
    PHOTO_CONSISTENCY, DEPTH_CONSISTENCY, PHOTO_DEPTH = 0, 1, 2

    def __init__(self, device=0):
        self._p = default_params()
        self._device = device
        self._p.method = PHOTO_CONSISTENCY        # the default argument of every method of the class (RPI.h:4519, 2545, 2745)
        self._ctx = None
        self._shape = None
        self._pending = {}
        self._res = None
        self._cam = None
        self._hg = None
        self.SSO = 0.0
        self.avResidual = 0.0          # never written on the occlusion-0 spherical path
        self.avPhotoResidual = 0.0
        self.avDepthResidual = 0.0
        self.nPyrLevels = self._p.n_levels

    # setters, RPI.h:224-269
    def setNumPyr(self, n):
        self._p.n_levels = int(n); self.nPyrLevels = int(n); self._drop()

    def setMinDepth(self, v):
        self._p.min_depth = float(v); self._drop()

    def setMaxDepth(self, v):
        self._p.max_depth = float(v); self._drop()

    def setGrayVariance(self, std):      # sets stdDevPhoto (RPI.h:242-245)
        self._p.std_photo = float(std); self._drop()

    def setDepthVariance(self, std):
        self._p.std_depth = float(std); self._drop()

    def setCameraMatrix(self, K):        # RPI.h:254 (3x3 camera matrix of the pinhole path)
        K = np.asarray(K, np.float32).reshape(3, 3)
        self._cam = (float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2]))
        if self._ctx is not None and self._p.projection == 1:
            self._ctx.set_camera(*self._cam)

    def setVisualization(self, viz):
        pass

    def useSaliency(self, flag):         # bUseSalientPixels branches are commented out upstream
        pass

    def _drop(self):
        if self._ctx is not None:
            self._ctx.close()
        self._ctx = None

    def _ensure(self, rgb):
        shape = rgb.shape[:2]
        if self._ctx is None or self._shape != shape:
            self._drop()
            self._ctx = Context(shape[0], shape[1], 2, 1, self._p, self._device)
            self._shape = shape
            if self._p.projection == 1:
                if self._cam is None:
                    raise RuntimeError("setCameraMatrix first (RPI.h:254)")
                self._ctx.set_camera(*self._cam)
            for slot, (r, d, role) in self._pending.items():
                if r.shape[:2] == shape:
                    self._ctx.set_frames(slot, r[None], d[None], [role])

    def setSourceFrame(self, rgb, depth):
        self._pending[0] = (rgb, depth, ROLE_SOURCE)
        self._ensure(rgb)
        self._ctx.set_frames(0, rgb[None], depth[None], [ROLE_SOURCE])

    def setTargetFrame(self, rgb, depth):
        self._pending[1] = (rgb, depth, ROLE_TARGET)
        self._ensure(rgb)
        self._ctx.set_frames(1, rgb[None], depth[None], [ROLE_TARGET])

    def _configure(self, method, occlusion, projection):
        """Re-create the context when the cost function, the occlusion variant or the registration changes."""
        if 0 not in self._pending or 1 not in self._pending:
            raise RuntimeError("setSourceFrame / setTargetFrame first")
        if self._ctx is None:
            self._ensure(self._pending[0][0])
        if method != self._p.method or occlusion != self._p.occlusion or projection != self._p.projection:
            self._p.method = int(method)
            self._p.occlusion = int(occlusion)
            if projection != self._p.projection:     # constants hard-coded in alignFrames / alignFrames360
                self._p.projection = int(projection)
                self._p.tol_residual = 1e-4 if projection == 1 else 1e-3      # RPI.h:4308 / 4594
                self._p.n_sensors_mask = 0 if projection == 1 else 8          # RPI.h:4537
            self._ctx.close(); self._ctx = None
            self._ensure(self._pending[0][0])

    def alignFrames(self, pose_guess=None, method=PHOTO_CONSISTENCY, occlusion=0):
        """The pinhole registration (RPI.h:4254-4512); needs setCameraMatrix."""
        if occlusion != 0:
            raise NotImplementedError("the pinhole occlusion variants (RPI.h:1107-2023) are not built")
        self._configure(method, 0, 1)
        guess = None if pose_guess is None else pose_to_colmajor(pose_guess)[None]
        self._res = self._ctx.register_pairs([0], [1], guess)[0]

    def errorPhotoICP(self, level, pose, method=PHOTO_CONSISTENCY):
        self._configure(method, 0, 1)
        e = self._ctx.eval_error_pinhole(0, 1, level, pose)
        return e["error"]

    def calcHessGrad(self, level, pose, method=PHOTO_CONSISTENCY):
        self._configure(method, 0, 1)
        H, g, nv = self._ctx.eval_hessgrad(0, 1, level, pose)
        self._hg = (H, g)

    def alignFrames360(self, pose_guess=None, method=PHOTO_CONSISTENCY, occlusion=0):
        if occlusion not in (0, 1, 2):
            raise ValueError("occlusion must be 0, 1 or 2 (RPI.h:4517)")
        self._configure(method, occlusion, 0)
        guess = None if pose_guess is None else pose_to_colmajor(pose_guess)[None]
        self._res = self._ctx.register_pairs([0], [1], guess)[0]
        self.SSO = float(self._res["sso"])

    def errorPhotoICP_sphere(self, level, pose, method=PHOTO_CONSISTENCY):
        self._configure(method, 0, 0)
        e2, n = self._ctx.eval_error(0, 1, level, pose)
        return float(np.sqrt(e2 / n)) if n else float("nan")

    def calcHessGrad_sphere(self, level, pose, method=PHOTO_CONSISTENCY):
        self._configure(method, 0, 0)
        H, g, nv = self._ctx.eval_hessgrad(0, 1, level, pose)
        self._hg = (H, g)
        self.SSO = nv / float((self._shape[0] >> level) * (self._shape[1] >> level))

    def getOptimalPose(self):
        return pose_from_colmajor(self._res["pose"])

    def getHessian(self):
        return np.array(self._res["hessian"], np.float32).reshape(6, 6)

    def getGradient(self):
        return np.array(self._res["gradient"], np.float32)
