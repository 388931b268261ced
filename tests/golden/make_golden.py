"""Generates the committed golden fixtures.  Run HERE (the container that has /root/reference and
python cv2); the GPU box only reads the fixtures.

  python tests/golden/make_golden.py

1. sample_pair.npz -- the reference's only data fixture (config #1 of BASELINE.json):
   samples/sphere_images_1.bin (target) and sphere_images_10.bin (source), parsed from the boost
   binary archive (third_party/cvSerialization/cvmat_serialization.h:22-54) and stitched to
   1920x320 RGB8 + depth u16 mm BY THE REFERENCE'S OWN CODE: Frame360::stitchSphericalImage / stitchImage
   (Frame360.h:386-405, 1099-1148) and Calib360 (extrinsics Calibration/Extrinsics/Rt_0N.txt, camera matrix
   Calib360.h:75-77) compiled from where they lie (oracle/ref_stitch_harness.cpp, glibc build).  The numpy
   restatement below is kept as an independent cross-check (printed).
2. cv2_vectors.npz -- OpenCV outputs that pin the third-party arithmetic the oracle restates:
   cvtColor(RGB2GRAY) on u8, convertTo-style scaling, pyrDown on f32 (python cv2 %s).
3. sample_pair_oracle.json -- the oracle's own result on the sample pair (poses, iterations,
   residual sums, counts) in both arithmetic modes: the regression vector the GPU run is
   compared with on the box (no reference-recorded output exists upstream).
"""
import json
import os
import struct
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def load_frame360(path):
    """8 x (RGB 240x320x3 u8, depth 240x320 u16) from a Frame360 .bin (Frame360.h:231-266)."""
    b = open(path, "rb").read()
    assert b[8:30] == b"serialization::archive"
    off = 45
    rgb, depth = [], []
    for k in range(16):
        cols, rows = struct.unpack_from("<ii", b, off)
        es, ty = struct.unpack_from("<QQ", b, off + 8)
        off += 24
        n = cols * rows * es
        a = np.frombuffer(b, np.uint8, n, off)
        off += n
        if k % 2 == 0:
            assert (es, ty) == (3, 16)
            rgb.append(a.reshape(rows, cols, 3).copy())
        else:
            assert (es, ty) == (2, 2)
            depth.append(a.view(np.uint16).reshape(rows, cols).copy())
    return rgb, depth


def stitch(rgb, depth):
    """stitchSphericalImage / stitchImage (Frame360.h:386-405, 1099-1148), float32 arithmetic."""
    f32 = np.float32
    size_h, size_w = rgb[0].shape[:2]                 # 240, 320
    W = size_h * 8
    H = int(W * 0.5 * 60.0 / 180)
    sphere_rgb = np.zeros((H, W, 3), np.uint8)
    sphere_d = np.zeros((H, W), np.uint16)
    fx = fy = f32(262.5); cx = f32(159.5); cy = f32(119.5)       # Calib360.h:75-77
    offsetPhi = f32(H // 2 - 0.5)
    offsetTheta = f32(-size_h * 15 // 2 + 0.5)
    angle_pixel = f32(2 * 3.14159265359 / W)
    for s in range(8):
        Rt = np.loadtxt(f"{REF}/Calibration/Extrinsics/Rt_0{s + 1}.txt").astype(np.float32)
        Rt_inv = np.linalg.inv(Rt.astype(np.float64)).astype(np.float32)
        rows = np.arange(H)
        cols = np.arange((7 - s) * size_h, (8 - s) * size_h)
        phi = ((offsetPhi - rows.astype(f32)) * angle_pixel).astype(f32)
        theta = ((cols.astype(f32) + offsetTheta) * angle_pixel).astype(f32)
        vp0 = np.sin(phi).astype(f32)[:, None] * np.ones_like(theta)[None, :]
        cphi = np.cos(phi).astype(f32)[:, None]
        vp1 = (cphi * np.sin(theta).astype(f32)[None, :]).astype(f32)
        vp2 = (cphi * np.cos(theta).astype(f32)[None, :]).astype(f32)
        R, t = Rt_inv[:3, :3], Rt_inv[:3, 3]
        pc = [((R[i, 0] * vp0 + R[i, 1] * vp1).astype(f32) + R[i, 2] * vp2).astype(f32) + t[i] for i in range(3)]
        u = (fx * pc[0] / pc[2] + cx).astype(f32)
        v = (fy * pc[1] / pc[2] + cy).astype(f32)
        ok = (u >= 0) & (u < size_w) & (v >= 0) & (v < size_h)
        ui = np.clip(u, 0, size_w - 1).astype(np.int64)       # at<>(v,u): float -> int truncation
        vi = np.clip(v, 0, size_h - 1).astype(np.int64)
        rr, cc = np.nonzero(ok)
        sphere_rgb[rows[rr], cols[cc]] = rgb[s][vi[rr, cc], ui[rr, cc]]
        a = ((u - cx) / fx).astype(f32); b = ((v - cy) / fy).astype(f32)
        scale = np.sqrt(((f32(1) + a * a).astype(f32) + b * b).astype(f32)).astype(f32)     # C++98: pow(float, 2), sqrt(float)
        dd = (depth[s][vi[rr, cc], ui[rr, cc]].astype(f32) * scale[rr, cc]).astype(f32)
        sphere_d[rows[rr], cols[cc]] = np.minimum(dd, 65535).astype(np.uint16)   # float -> ushort truncation
    return sphere_rgb, sphere_d


def main():
    import cv2
    from oracle import orc, refbind
    # ---- 1. sample pair
    out = {}
    for name, fn in (("trg", "sphere_images_1.bin"), ("src", "sphere_images_10.bin")):
        rgb, depth = load_frame360(f"{REF}/samples/{fn}")
        srgb, sd, _ = refbind.stitch(np.stack(rgb), np.stack(depth), pinned=False)     # the reference's own stitch
        n_rgb, n_d = stitch(rgb, depth)
        out[name + "_rgb"] = srgb
        out[name + "_depth"] = sd
        print(name, srgb.shape, sd.shape, "valid depth %.3f" % np.mean((sd > 300) & (sd < 6000)),
              "| numpy restatement differs on", int((n_rgb != srgb).any(-1).sum()), "rgb /", int((n_d != sd).sum()), "depth pixels")
    np.savez_compressed(os.path.join(HERE, "sample_pair.npz"), **out)

    # ---- 1b. raw sensor images of sphere_images_1.bin (ingest fixture, tests/test_ingest.py)
    import hashlib
    path = f"{REF}/samples/sphere_images_1.bin"
    b = open(path, "rb").read()
    rgb, depth = load_frame360(path)
    body = 16 * 24 + sum(a.nbytes for a in rgb) + sum(a.nbytes for a in depth)
    Rt = np.stack([np.loadtxt(f"{REF}/Calibration/Extrinsics/Rt_0{s + 1}.txt") for s in range(8)])
    np.savez_compressed(os.path.join(HERE, "frame360_raw_1.npz"), rgb=np.stack(rgb), depth=np.stack(depth),
                        preamble=np.frombuffer(b[:45], np.uint8), tail=np.frombuffer(b[45 + body:], np.uint8), Rt=Rt,
                        sha256=np.array(hashlib.sha256(b).hexdigest()))

    # ---- 2. cv2 vectors
    rng = np.random.default_rng(360)
    rgb = rng.integers(0, 256, (48, 64, 3), dtype=np.uint8)
    rgb[0, 0] = 255
    gray = cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY)
    gray[0, 0] = 255
    gray_f = cv2.normalize(gray, None, alpha=1, beta=0, norm_type=cv2.NORM_INF, dtype=cv2.CV_32F)   # convertTo(CV_32F, 1/255)
    f = rng.random((64, 96), dtype=np.float32)
    pd = cv2.pyrDown(f, dstsize=(48, 32))
    f2 = out["trg_rgb"][..., 1].astype(np.float32)[:, :256] / np.float32(255)
    pd2 = cv2.pyrDown(f2, dstsize=(128, 160))
    np.savez_compressed(os.path.join(HERE, "cv2_vectors.npz"), rgb=rgb, gray=cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY),
                        gray_u8=gray, gray_f=gray_f, pyr_in=f, pyr_out=pd, pyr_in2=f2, pyr_out2=pd2,
                        cv2_version=np.array(cv2.__version__))

    # ---- 3. oracle result on the sample pair (regression vector)
    P = orc.default_params(n_levels=4)
    res = {}
    for mode, mname in ((orc.MATH_PINNED, "pinned"), (orc.MATH_LIBM, "libm")):
        orc.set_math(mode)
        trg = orc.Frame(out["trg_rgb"], out["trg_depth"], P, True)
        src = orc.Frame(out["src_rgb"], out["src_depth"], P, False)
        r, tr = orc.align(src, trg, None, P, trace=True)
        res[mname] = dict(pose=[float(x) for x in r.pose], iters=list(r.iters)[:4], passes=list(r.passes)[:4],
                          final_err2=r.final_err2, final_n_valid=r.final_n_valid, final_error=r.final_error,
                          status=r.status, sso=float(r.sso), n_visible=r.n_visible,
                          hessian=[float(x) for x in r.hessian], gradient=[float(x) for x in r.gradient],
                          trace=[dict(level=t.level, it=t.it, accepted=t.accepted, err2=t.err2, n_valid=t.n_valid,
                                      n_visible=t.n_visible, used=t.used) for t in tr if t.used])
        print(mname, res[mname]["iters"], res[mname]["final_error"], res[mname]["final_n_valid"])
    orc.set_math(orc.MATH_PINNED)
    json.dump(res, open(os.path.join(HERE, "sample_pair_oracle.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
