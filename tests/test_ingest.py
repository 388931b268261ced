"""Frame360 ingest: the .bin parser (Frame360::loadFrame, Frame360.h:231-266) and
stitchSphericalImage (Frame360.h:386-405, 1099-1148).

Pinned against THE REFERENCE'S OWN CODE: oracle/_ref/librpi_ref_stitch[_pinned].so holds Calib360.h (whole) and the two
Frame360 member functions (verbatim, cut out of Frame360.h at build time) compiled against the third-party stand-ins
(oracle/ref_stitch_harness.cpp).  tests/golden/reference_stitch.json records its output on the raw sensor images of the
reference's samples/sphere_images_1.bin (fixture tests/golden/frame360_raw_1.npz: the 8 RGB + 8 depth sensor images,
the archive's 45-byte preamble and 24-byte tail, the extrinsics Calibration/Extrinsics/Rt_0N.txt, the sha256 of the
original file); tests/golden/sample_pair.npz IS that output (glibc build) for both sample files.
The oracle restatement and the CUDA stitch must equal it bit for bit (u8 / u16 outputs).
"""
import hashlib
import json
import os
import struct
import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


@pytest.fixture(scope="module")
def rec():
    with open(os.path.join(GOLD, "reference_stitch.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def raw(rec):
    z = np.load(os.path.join(GOLD, "frame360_raw_1.npz"))
    # Rt_inv as Calib360::loadExtrinsicCalibration computed it in the recorded reference run (Calib360.h:128-129)
    Rt_inv = np.array(rec["Rt_inv"], np.float32).reshape(8, 4, 4)
    return dict(rgb=z["rgb"], depth=z["depth"], preamble=z["preamble"].tobytes(), tail=z["tail"].tobytes(),
                sha256=str(z["sha256"]), Rt_inv=Rt_inv, Rt=z["Rt"])


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


def test_oracle_stitch_equals_recorded_reference(orc, raw, rec):
    """The oracle's stitch == the reference's own stitch code on the reference's own sample frame: every byte of
    the sphere RGB and depth images, glibc and pinned trig builds; the recorded camera matrix is Calib360's."""
    assert rec["camera"] == [262.5, 262.5, 159.5, 119.5]                            # Calib360.h:75-77
    np.testing.assert_allclose(raw["Rt_inv"], [np.linalg.inv(raw["Rt"][s]) for s in range(8)], atol=2e-7)
    want = np.load(os.path.join(GOLD, "sample_pair.npz"))
    try:
        for mode, key in ((orc.MATH_LIBM, "libm"), (orc.MATH_PINNED, "pinned")):
            orc.set_math(mode)
            rgb, d = orc.stitch(raw["rgb"], raw["depth"], raw["Rt_inv"])
            assert rgb.shape == (rec[key]["rows"], rec[key]["cols"], 3) == (320, 1920, 3)
            assert _digest(rgb) == rec[key]["rgb_sha"] and _digest(d) == rec[key]["depth_sha"], key
            assert int((d > 0).sum()) == rec[key]["valid_depth"]
            if key == "libm":                                                      # config #1's target frame is this image
                assert np.array_equal(rgb, want["trg_rgb"]) and np.array_equal(d, want["trg_depth"])
    finally:
        orc.set_math(orc.MATH_PINNED)


def _random_rig_case(seed):
    """Seeded random QVGA sensor images (random colours, depths with holes) and the reference's rig perturbed by
    ~1 degree / 3 cm per sensor: other pixel-boundary ties and misses than the recorded frame."""
    rng = np.random.default_rng(7000 + seed)
    h, w = 240, 320
    rgb = rng.integers(0, 256, (8, h, w, 3), dtype=np.uint8)
    depth = rng.integers(0, 9000, (8, h, w)).astype(np.uint16)
    depth[rng.random(depth.shape) < 0.1] = 0
    Rt = np.load(os.path.join(GOLD, "frame360_raw_1.npz"))["Rt"]
    Rt_inv = []
    for s in range(8):
        wv = rng.normal(0, 0.02, 3)
        K = np.array([[0, -wv[2], wv[1]], [wv[2], 0, -wv[0]], [-wv[1], wv[0], 0]])
        T = Rt[s].astype(np.float64).copy()
        T[:3, :3] = (np.eye(3) + K) @ T[:3, :3]
        T[:3, 3] += rng.normal(0, 0.03, 3)
        Rt_inv.append(np.linalg.inv(T))
    return rgb, depth, np.array(Rt_inv, np.float32), h


@pytest.mark.parametrize("seed", range(6))
def test_oracle_stitch_equals_live_reference_on_random_rigs(orc, seed):
    """Beyond the recorded frame: random sensor images and rigs through the compiled reference HERE (container only)
    and through the oracle, both trig builds, every byte."""
    from oracle import refbind
    if not refbind.stitch_available():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    rgb, depth, Rt_inv, h = _random_rig_case(seed)
    try:
        for pinned in (False, True):
            r_rgb, r_d, used = refbind.stitch(rgb, depth, Rt_inv=Rt_inv, pinned=pinned)
            assert np.array_equal(used, Rt_inv)
            orc.set_math(orc.MATH_PINNED if pinned else orc.MATH_LIBM)
            # the reference's geometry is hard-wired to its QVGA camera matrix (Calib360.h:75-77)
            o_rgb, o_d = orc.stitch(rgb, depth, Rt_inv)
            assert np.array_equal(r_rgb, o_rgb), (seed, pinned, int((r_rgb != o_rgb).any(-1).sum()))
            assert np.array_equal(r_d, o_d), (seed, pinned, int((r_d != o_d).sum()))
    finally:
        orc.set_math(orc.MATH_PINNED)


def test_live_reference_stitch_matches_recording(raw, rec):
    """Guards against a stale reference_stitch.json (container only)."""
    from oracle import refbind
    if not refbind.stitch_available() or not os.path.isdir("/root/reference/Calibration/Extrinsics"):
        pytest.skip("needs /root/reference")
    for pinned, key in ((False, "libm"), (True, "pinned")):
        rgb, d, used = refbind.stitch(raw["rgb"], raw["depth"], pinned=pinned)           # Calib360::loadExtrinsicCalibration
        assert _digest(rgb) == rec[key]["rgb_sha"] and _digest(d) == rec[key]["depth_sha"]
        assert np.array_equal(used, raw["Rt_inv"])


@pytest.mark.gpu
def test_cuda_stitch_bit_exact_and_feeds_the_path(orc, r360, raw, rec):
    rows, cols = r360.native.sphere_shape(240)
    assert (rows, cols) == (320, 1920)
    rig = r360.native.make_rig(raw["Rt_inv"])
    L = 4
    ctx = r360.Context(rows, cols, 3, 1, r360.default_params(n_levels=L))
    try:
        srgb, sdep = ctx.stitch_frames(rig, 1, raw["rgb"][None], raw["depth"][None], [r360.ROLE_TARGET])
        # k_stitch == the reference's own stitch code (pinned-trig build) as recorded: every byte
        assert _digest(srgb[0]) == rec["pinned"]["rgb_sha"] and _digest(sdep[0]) == rec["pinned"]["depth_sha"]
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


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(6))
def test_cuda_stitch_random_rigs(orc, r360, seed):
    """k_stitch == the oracle (== the compiled reference, test_oracle_stitch_equals_live_reference_on_random_rigs) on
    random sensor images and rigs, several frames per call, every byte."""
    rgb, depth, Rt_inv, h = _random_rig_case(seed)
    rows, cols = r360.native.sphere_shape(h)
    rig = r360.native.make_rig(Rt_inv, sensor_rows=h, sensor_cols=rgb.shape[2])
    o_rgb, o_d = orc.stitch(rgb, depth, Rt_inv)
    L = 1 if rows % 2 else 2
    if cols % (1 << L):
        L = 1
    ctx = r360.Context(rows, cols, 2, 1, r360.default_params(n_levels=L, n_sensors_mask=0))
    try:
        srgb, sdep = ctx.stitch_frames(rig, 0, np.stack([rgb, rgb[::-1]]), np.stack([depth, depth[::-1]]))
        assert np.array_equal(srgb[0], o_rgb) and np.array_equal(sdep[0], o_d)
        o_rgb2, o_d2 = orc.stitch(rgb[::-1], depth[::-1], Rt_inv)
        assert np.array_equal(srgb[1], o_rgb2) and np.array_equal(sdep[1], o_d2)
    finally:
        ctx.close()
