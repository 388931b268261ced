"""Host drivers that feed batches of sphere pairs to the registration path -- Python mirror of
include/Drivers360_b200.hpp (what the reference's callers do around alignFrames360):
Registration/OdometryRGBD360.cpp:137-139, 185-193 and include/LoopClosure360.h:119-126, 291-321."""
import numpy as np

from .native import Context, default_params, pose_to_colmajor, PHOTO_DEPTH


def rot_offset():
    """rotOffset: 157.5 deg about x (OdometryRGBD360.cpp:137-139), float32 4x4."""
    a = np.float32(157.5 * 3.14159265359 / 180)
    R = np.eye(4, dtype=np.float32)
    R[1, 1] = R[2, 2] = np.cos(a)
    R[1, 2] = np.sin(a)
    R[2, 1] = -R[1, 2]
    return R


def _inv(T):
    T = np.asarray(T, np.float32)
    I = np.eye(4, dtype=np.float32)
    I[:3, :3] = T[:3, :3].T
    I[:3, 3] = -(T[:3, :3].T @ T[:3, 3])
    return I


def to_sphere_frame(robot_pose):
    """Initial guess in the sphere-image frame: rotOffset * relativePose * rotOffset^-1 (LoopClosure360.h:311)."""
    R = rot_offset()
    return (R @ np.asarray(robot_pose, np.float32) @ _inv(R)).astype(np.float32)


def to_robot_frame(sphere_pose):
    """rotOffset^-1 * getOptimalPose() * rotOffset (OdometryRGBD360.cpp:193)."""
    R = rot_offset()
    return (_inv(R) @ np.asarray(sphere_pose, np.float32) @ R).astype(np.float32)


def loop_candidates(kf_poses, new_id, max_dist=5.0):
    """(compare id, new id) for every keyframe closer than max_dist to `new_id` (LoopClosure360.h:289-293)."""
    out = []
    for k, P in enumerate(kf_poses):
        if k == new_id:
            continue
        rel = _inv(P) @ np.asarray(kf_poses[new_id], np.float32)
        if np.linalg.norm(rel[:3, 3]) < max_dist:
            out.append((k, new_id))
    return out


class BatchRegistrar:
    """Frames resident on one GPU (both roles) + batched pair lists over them."""

    def __init__(self, rows, cols, max_frames, max_pairs, n_levels=5, std_photo=3.0 / 255, device=0):
        self.params = default_params(n_levels=n_levels, std_photo=np.float32(std_photo), method=PHOTO_DEPTH)
        self.ctx = Context(rows, cols, max_frames, max_pairs, self.params, device=device)

    def close(self):
        self.ctx.close()

    def set_frames(self, first, rgb, depth_mm):
        self.ctx.set_frames(first, rgb, depth_mm, None)

    def set_frames_from_sensors(self, rig, first, sensor_rgb, sensor_depth_mm):
        self.ctx.stitch_frames(rig, first, sensor_rgb, sensor_depth_mm, None, want_sphere=False)

    def _edges(self, src, trg, guesses):
        res = self.ctx.register_pairs(src, trg, guesses)
        return [dict(source=int(s), target=int(t),
                     relativePose=to_robot_frame(np.array(r["pose"], np.float32).reshape(4, 4).T),
                     informationMatrix=np.array(r["hessian"], np.float32).reshape(6, 6),
                     SSO=float(r["sso"]), status=int(r["status"]), iterations=list(r["iters"][:self.params.n_levels]))
                for s, t, r in zip(src, trg, res)]

    def odometry(self, first, n, robot_guesses=None):
        """Edges (k+1 -> k) of frames [first, first + n).  robot_guesses: optional n-1 robot-frame guesses.  The
        reference seeds pair k with the PREVIOUS pair's result (rigidTransf_dense, OdometryRGBD360.cpp:191), a serial
        dependency a batched call cannot have: the default here is Identity (the alternative the reference's own
        driver offers, MethodsRegisterRGBD360.cpp:448); callers with a motion prior pass it explicitly."""
        trg = np.arange(first, first + n - 1, dtype=np.int32)
        g = None
        if robot_guesses is not None:
            if len(robot_guesses) != n - 1:
                raise ValueError(f"{len(robot_guesses)} guesses for {n - 1} pairs")
            g = np.stack([pose_to_colmajor(to_sphere_frame(G)) for G in robot_guesses])
        return self._edges(trg + 1, trg, g)

    def loop_closures(self, candidates, robot_guesses):
        src = np.array([c[0] for c in candidates], np.int32)
        trg = np.array([c[1] for c in candidates], np.int32)
        g = np.stack([pose_to_colmajor(to_sphere_frame(G)) for G in robot_guesses]) if len(candidates) else None
        return self._edges(src, trg, g)
