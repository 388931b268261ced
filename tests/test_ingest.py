"""Frame360 ingest: the .bin parser (Frame360::loadFrame, Frame360.h:231-266) and
stitchSphericalImage (Frame360.h:386-405, 1099-1148).

Fixture tests/golden/frame360_raw_1.npz = the 8 RGB + 8 depth sensor images of the reference's
samples/sphere_images_1.bin, the archive's 45-byte preamble and 24-byte tail, the extrinsics
Calibration/Extrinsics/Rt_0N.txt and the sha256 of the original file (tests/golden/make_golden.py).
Frame360.h itself cannot be compiled here (PCL segmentation, boost serialization, ...), so the stitch
oracle is a restatement checked against the independent numpy restatement that produced
tests/golden/sample_pair.npz; the CUDA stitch must equal the oracle bit for bit (u8 / u16 outputs).
"""
import hashlib
import os
import struct
import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def raw():
    z = np.load(os.path.join(GOLD, "frame360_raw_1.npz"))
    Rt_inv = np.stack([np.linalg.inv(z["Rt"][s].astype(np.float32).astype(np.float64)).astype(np.float32) for s in range(8)])
    return dict(rgb=z["rgb"], depth=z["depth"], preamble=z["preamble"].tobytes(), tail=z["tail"].tobytes(),
                sha256=str(z["sha256"]), Rt_inv=Rt_inv)


def _archive(raw):
    """Re-assembles the original .bin: preamble, 16 x {int32 cols, int32 rows, uint64 elemSize,
    uint64 cvType, pixels} (cvmat_serialization.h:22-37), tail (the timestamp Mat)."""
    out = [raw["preamble"]]
    for s in range(8):
        h, w = raw["depth"].shape[1:]
        out.append(struct.pack("<iiQQ", w, h, 3, 16) + raw["rgb"][s].tobytes())
        out.append(struct.pack("<iiQQ", w, h, 2, 2) + raw["depth"][s].tobytes())
    out.append(raw["tail"])
    return b"".join(out)


def test_frame360_parse_roundtrip(r360, raw):
    data = _archive(raw)
    assert hashlib.sha256(data).hexdigest() == raw["sha256"]          # byte-identical to the reference's sample file
    rgb, dep = r360.native.frame360_parse(data)
    assert rgb.shape == (8, 240, 320, 3) and dep.shape == (8, 240, 320)
    assert np.array_equal(rgb, raw["rgb"]) and np.array_equal(dep, raw["depth"])
    for bad in (data[:1000], b"x" * 5000, data[:45] + b"\xff" * 4000):
        with pytest.raises(r360.R360Error):
            r360.native.frame360_parse(bad)


def test_oracle_stitch_matches_numpy_restatement(orc, raw):
    """Two independent restatements of stitchImage (this oracle, C++ float; make_golden.stitch, numpy):
    identical up to the few pixels whose (u, v) fall within an ulp of a sensor-pixel boundary."""
    want = np.load(os.path.join(GOLD, "sample_pair.npz"))
    try:
        for mode in (orc.MATH_LIBM, orc.MATH_PINNED):
            orc.set_math(mode)
            rgb, d = orc.stitch(raw["rgb"], raw["depth"], raw["Rt_inv"])
            assert rgb.shape == (320, 1920, 3)
            assert int((rgb != want["trg_rgb"]).any(-1).sum()) <= 10
            assert int((d != want["trg_depth"]).sum()) <= 12
    finally:
        orc.set_math(orc.MATH_PINNED)


@pytest.mark.gpu
def test_cuda_stitch_bit_exact_and_feeds_the_path(orc, r360, raw):
    rows, cols = r360.native.sphere_shape(240)
    assert (rows, cols) == (320, 1920)
    rig = r360.native.make_rig(raw["Rt_inv"])
    L = 4
    ctx = r360.Context(rows, cols, 3, 1, r360.default_params(n_levels=L))
    try:
        srgb, sdep = ctx.stitch_frames(rig, 1, raw["rgb"][None], raw["depth"][None], [r360.ROLE_TARGET])
        o_rgb, o_d = orc.stitch(raw["rgb"], raw["depth"], raw["Rt_inv"])            # PINNED arithmetic
        assert np.array_equal(srgb[0], o_rgb)                                        # u8: bit-exact
        assert np.array_equal(sdep[0], o_d)                                          # u16: bit-exact
        # the stitched frame sits in slot 1 as a target: its pyramid == the oracle's on the same sphere
        P = orc.default_params(n_levels=L)
        trg = orc.Frame(o_rgb, o_d, P, True)
        for level in range(L):
            g = ctx.dump_level(1, level); o = trg.level(level)
            for k in o:
                assert np.array_equal(g[k].view(np.int32), o[k].view(np.int32)), (level, k)
        # config #1 from the raw sensor images: source = the stitched sphere_images_10 fixture
        z = np.load(os.path.join(GOLD, "sample_pair.npz"))
        ctx.set_frames(0, z["src_rgb"][None], z["src_depth"][None], [r360.ROLE_SOURCE])
        res = ctx.register_pairs([0], [1])[0]
        src = orc.Frame(z["src_rgb"], z["src_depth"], P, False)
        o = orc.align(src, trg, None, P)
        assert list(res["iters"][:L]) == list(o.iters)[:L]
        assert np.allclose(np.array(res["pose"]), np.array(o.pose), atol=1e-4)
        # a wrong geometry is refused loudly
        bad = r360.Context(256, 512, 1, 1, r360.default_params(n_levels=2))
        with pytest.raises(r360.R360Error):
            bad.stitch_frames(rig, 0, raw["rgb"][None], raw["depth"][None])
        bad.close()
    finally:
        ctx.close()
