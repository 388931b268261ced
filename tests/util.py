"""Shared helpers for the parity tests."""
import numpy as np


def pose_err(T, G):
    """(rotation angle [rad], translation distance [m]) between two 4x4 poses."""
    T = np.asarray(T, np.float64); G = np.asarray(G, np.float64)
    dR = T[:3, :3] @ G[:3, :3].T
    # atan2 form: arccos((tr-1)/2) loses half the digits near 0 (float32 poses -> 4e-4 rad noise)
    v = 0.5 * np.array([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]])
    ang = float(np.arctan2(np.linalg.norm(v), (np.trace(dR) - 1) / 2))
    return ang, float(np.linalg.norm(T[:3, 3] - G[:3, 3]))


def small_pose(rx=0.0, ry=0.0, rz=0.0, tx=0.0, ty=0.0, tz=0.0):
    def R(axis, a):
        c, s = np.cos(a), np.sin(a)
        M = np.eye(3)
        i, j = [(1, 2), (2, 0), (0, 1)][axis]
        M[i, i] = c; M[i, j] = -s; M[j, i] = s; M[j, j] = c
        return M
    T = np.eye(4)
    T[:3, :3] = R(0, rx) @ R(1, ry) @ R(2, rz)
    T[:3, 3] = [tx, ty, tz]
    return T.astype(np.float32)


def upper21(H):
    H = np.asarray(H).reshape(6, 6)
    return np.array([H[a, b] for a in range(6) for b in range(a, 6)])
