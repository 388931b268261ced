"""rgbd360_b200 -- B200-native batched spherical dense RGB-D registration.

Python binding (ctypes) of the C ABI in include/r360.h.  The product path is the CUDA shared
library rgbd360_b200/librgbd360_b200.so; importing the package works without it, every call
fails loudly if it is missing.  Nothing here imports the CPU oracle.
"""
from .native import (Context, Params, Result, IterRecord, default_params, lib, build_native,
                     PHOTO_CONSISTENCY, DEPTH_CONSISTENCY, PHOTO_DEPTH, ROLE_SOURCE, ROLE_TARGET,
                     ROLE_BOTH, R360Error, pose_to_colmajor, pose_from_colmajor, synth_gt_pose, pinhole_params)
from .register import RegisterPhotoICP

__all__ = ["Context", "Params", "Result", "IterRecord", "default_params", "lib", "build_native",
           "PHOTO_CONSISTENCY", "DEPTH_CONSISTENCY", "PHOTO_DEPTH", "ROLE_SOURCE", "ROLE_TARGET",
           "ROLE_BOTH", "R360Error", "RegisterPhotoICP", "pose_to_colmajor", "pose_from_colmajor",
           "synth_gt_pose", "pinhole_params"]
