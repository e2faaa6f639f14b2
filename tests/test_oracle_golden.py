"""Pin the CPU oracle against every fixture the reference offers for this path.
The reference has no recorded outputs, so these are: the survey's independent numpy restatement
(Appendix B CRCs), ST's generated tables in network.c / network_data.c, and the quantisation
constants duplicated in yoloface.c / tflite_prediction.py."""
import zlib

import numpy as np
import pytest

from oracle_lib import vector_a, vector_b

# SURVEY.md Appendix B, vector A: CRC32 of every op's output tensor
APPENDIX_B = """op0:c1d54981 op1:a1dae5ea op2:2e2afcc8 op3:641eb8fb op4:f3af3943 op5:28ff72b6 op6:8ebe3c08 op7:769f8485
op8:ac0b0e76 op9:e4b797ff op10:0cb53d2e op11:5a60e438 op12:04e64f84 op13:febdc6d5 op14:d25f4134 op15:f52b8ac3
op16:0ad948ac op17:2106b104 op18:9684933d op19:9abfc23b op20:d690ce33 op21:473bf608 op22:368cbbc1 op23:cf821bc5
op24:dc4dd9bc op25:5a35d042 op26:aa3e7fe7 op27:4122d713 op28:cf4bf9ec op29:818a5aac op30:04234fff op31:2232b4f6
op32:5134a25f op33:014e774b op34:564e9d5f op35:efbf0cd6 op36:a7946f71 op37:4d0730b7 op38:8b96d1c6 op39:d04cc252
op40:c8aaf2b9 op41:73409191 op42:2aafd862 op43:1abdbdb6 op44:fffb7520 op45:37d714b8 op46:9a81cd69 op47:d64c521c
op48:95145340 op49:28ebfaf9 op50:71e4b639 op51:e52b0eee op52:211e5be5 op53:f50a1d75"""


def crc(a):
    return "%08x" % zlib.crc32(np.ascontiguousarray(a).tobytes())


def test_model_structure(oracle):
    # network_generate_report.txt:15-21, SURVEY.md section 0
    assert oracle.num_ops == 54 and oracle.num_tensors == 104
    counts = {}
    for i in range(54):
        c = oracle.op(i)["opcode"]; counts[c] = counts.get(c, 0) + 1
    assert counts == {34: 3, 3: 17, 4: 7, 98: 17, 17: 2, 0: 3, 2: 2, 114: 3}
    tin, tout = oracle.tensor(0), oracle.tensor(100)
    assert tin["shape"] == [1, 56, 56, 3] and tin["zp"] == [-128]
    assert abs(tin["scale"][0] - 0.003921568859368563) < 1e-12
    # yoloface.c:116 / tflite_prediction.py:44 / network_generate_report.txt:17
    assert tout["shape"] == [1, 7, 7, 18] and tout["zp"] == [-15]
    assert np.float32(tout["scale"][0]) == np.float32(0.14218327403068542)


def test_vector_a_per_op_crcs(oracle):
    a = vector_a()
    assert crc(a) == "53e37af1"
    head, outs = oracle.run(a, dump=True)
    exp = dict(x.split(":") for x in APPENDIX_B.split())
    bad = [i for i in range(54) if crc(outs[i]) != exp["op%d" % i]]
    assert bad == []
    assert crc(head) == "f50a1d75" and int(head.astype(np.int64).sum()) == -10068
    assert head[0, 0].tolist() == [-8, -14, -15, -17, -39, 64, -4, -8, -15, -17, -51, 49, -7, -7, -16, -17, -74, 36]


def test_vector_b(oracle):
    b = vector_b()
    assert crc(b) == "c46eb5fc"
    head = oracle.run(b)
    assert crc(head) == "6f698aa4" and int(head.astype(np.int64).sum()) == -10374
    assert head[0, 0].tolist() == [-15, -15, -16, -18, -45, 63, -14, -10, -14, -17, -73, 45, -13, -9, -15, -18, -80, 34]


def test_committed_pins(oracle, golden):
    assert np.array_equal(oracle.run(vector_a()), golden["head_a"])
    assert np.array_equal(oracle.run(vector_b()), golden["head_b"])
    assert np.array_equal(oracle.run_batch(golden["images"], threads=4), golden["heads_images"])


def test_st_leaky_luts_negative_control(oracle, golden):
    """ST's 17 activation LUTs (network.c:2218..2902) are float-rounded; TFLite's fixed-point
    LEAKY_RELU differs from them in exactly 271 of 4,352 entries, always by 1 LSB."""
    ops, luts = golden["st_lut_ops"], golden["st_luts"]
    leaky_ops = [i for i in range(54) if oracle.op(i)["opcode"] == 98]
    assert list(ops) == leaky_ops
    ndiff = 0
    for op, st in zip(ops, luts):
        mine = oracle.leaky_lut(int(op)).astype(np.int32)
        d = np.abs(mine - st.astype(np.int32))
        assert d.max() <= 1
        ndiff += int((d != 0).sum())
        # and ST's table is the float32 round-half-even formula (so the parse is right)
        o = oracle.op(int(op)); ti, to = oracle.tensor(o["inputs"][0]), oracle.tensor(o["output"])
        q = np.arange(-128, 128, dtype=np.float32)
        v = (q - np.float32(ti["zp"][0])) * np.float32(ti["scale"][0])
        v = np.where(q >= ti["zp"][0], v, v * np.float32(0.1)) / np.float32(to["scale"][0])
        ref = np.clip(np.rint(v) + to["zp"][0], -128, 127).astype(np.int32)
        assert np.array_equal(ref, st.astype(np.int32))
    assert ndiff == 271


def test_st_weight_blob_identity(oracle, golden):
    """network_data.c's blob is byte-identical to the .tflite conv buffers at the offsets that
    network.c:3117-3263 binds (weights OHWI / 1HWC int8, then int32 bias)."""
    blob = golden["st_blob"].tobytes()
    assert len(blob) == 11304
    convs = [i for i in range(54) if oracle.op(i)["opcode"] in (3, 4)]
    assert list(golden["st_blob_ids"]) == convs
    for (woff, boff), op in zip(golden["st_blob_offsets"], convs):
        o = oracle.op(op)
        w, b = oracle.tensor(o["inputs"][1])["data"], oracle.tensor(o["inputs"][2])["data"]
        assert blob[woff:woff + len(w)] == w, op
        assert blob[boff:boff + len(b)] == b, op


def test_st_intq_tables_agree_with_flatbuffer(oracle, golden):
    """Every activation scale/zero-point appears twice: in the .tflite and in network.c:663-1341."""
    fb = {}
    for t in range(oracle.num_tensors):
        ti = oracle.tensor(t)
        if ti["type"] == 9 and len(ti["scale"]) == 1 and not ti["data"]:
            fb.setdefault((np.float32(ti["scale"][0]).tobytes(), int(ti["zp"][0])), t)
    hits = 0
    for name, sc, zp in zip(golden["st_intq_names"], golden["st_intq_scale"], golden["st_intq_zp"]):
        if str(name).endswith("_output") or str(name).startswith("Input"):
            assert (np.float32(sc).tobytes(), int(zp)) in fb, name
            hits += 1
    assert hits >= 30


@pytest.mark.parametrize("seed", [0, 1])
def test_fixed_point_primitives_match_python_bigint(oracle, seed):
    rng = np.random.default_rng(seed)
    lib = oracle.lib
    xs = list(rng.integers(-2**31, 2**31, 2000)) + [0, 1, -1, 2**31 - 1, -2**31, 2**30, -2**30]
    ms = list(rng.integers(2**30, 2**31, 2000)) + [2**30, 2**31 - 1, -2**31, 0, 1, -1, 1073741824]
    for x, m in zip(xs, ms[:len(xs)]):
        x, m = int(x), int(m)
        if x == m == -2**31:
            exp = 2**31 - 1
        else:
            ab = x * m; nudge = (1 << 30) if ab >= 0 else 1 - (1 << 30)
            t = ab + nudge; exp = abs(t) // (1 << 31) * (1 if t >= 0 else -1)   # trunc toward zero
        assert lib.yfo_srdhm(x, m) == exp
        # the form the CUDA epilogue uses: floor((ab + 2^30) / 2^31)
        if not (x == m == -2**31):
            assert exp == (x * m + (1 << 30)) >> 31
    for x in xs:
        x = int(x)
        for e in (0, 1, 5, 9, 12, 31):
            mask = (1 << e) - 1; rem = x & mask; thr = (mask >> 1) + (1 if x < 0 else 0)
            assert lib.yfo_rdivpot(x, e) == (x >> e) + (1 if rem > thr else 0)


def test_kernel_requant_form_is_exact(oracle):
    """The 32-bit form the CUDA epilogues evaluate (csrc/yf_requant.cuh):
        a = (acc + bias') << 9;  u = mulhi(a, m) + 2^7 + 256 * c2p - 256 * [a < 0];  idx = clamp(u >> (8 + e), 0, 255)
    equals clamp(MultiplyByQuantizedMultiplier(acc + bias', m, -e) + zp_out + 128, 0, 255) of the oracle for every
    |acc + bias'| < 2^22, 2^30 < m < 2^31, 1 <= e <= 13 -- the ranges csrc/yf_plan.cc::epi_lean_words admits."""
    rng = np.random.default_rng(2026)
    lib = oracle.lib
    edge = [0, 1, -1, 2, -2, 2**22 - 1, -(2**22 - 1), 255, -255, 4096, -4096]
    n = 0
    for e in range(1, 14):
        for m in [2**30 + 1, 2**31 - 1] + [int(v) for v in rng.integers(2**30 + 1, 2**31, 12)]:
            for zp in (-128, -15, 0, 127, int(rng.integers(-128, 128))):
                c2p = (1 << (e - 1)) + ((zp + 128) << e)
                kc = 128 + 256 * c2p
                # values around every rounding boundary of the two-stage rounding as well as random ones
                xs = edge + [int(v) for v in rng.integers(-(2**22) + 1, 2**22, 60)]
                for t in rng.integers(-300, 300, 6):                       # x with SRDHM(x, m) ~ t * 2^e +- half
                    base = (int(t) << e) + (1 << (e - 1))
                    x0 = (base << 31) // m
                    xs += [x0 - 1, x0, x0 + 1, -x0 - 1, -x0, -x0 + 1]
                for x in xs:
                    if abs(x) >= 2**22:
                        continue
                    a = x * 512
                    u = ((a * m) >> 32) + kc - (256 if a < 0 else 0)          # mulhi = floor(a * m / 2^32)
                    got = min(255, max(0, u >> (8 + e)))
                    want = min(255, max(0, lib.yfo_mbqm(x, m, -e) + zp + 128))
                    assert got == want, (x, m, e, zp)
                    assert -2**31 <= u < 2**31 and -2**31 <= a < 2**31
                    n += 1
    assert n > 50000


def test_quantize_multiplier_known_answers(oracle):
    import ctypes as C
    m, s = C.c_int32(), C.c_int()
    lib = oracle.lib
    lib.yfo_quantize_multiplier(0.5, m, s); assert (m.value, s.value) == (1 << 30, 0)
    lib.yfo_quantize_multiplier(1.0, m, s); assert (m.value, s.value) == (1 << 30, 1)
    lib.yfo_quantize_multiplier(0.0, m, s); assert (m.value, s.value) == (0, 0)
    lib.yfo_quantize_multiplier(0.75, m, s); assert (m.value, s.value) == (3 << 29, 0)
    # LEAKY_RELU op 2 identity / alpha multipliers
    o = oracle.op(2); ti, to = oracle.tensor(o["inputs"][0]), oracle.tensor(o["output"])
    ident = float(np.float32(ti["scale"][0]) / np.float32(to["scale"][0]))
    alpha = float(np.float32(ti["scale"][0]) * np.float32(0.10000000149011612) / np.float32(to["scale"][0]))
    lib.yfo_quantize_multiplier(ident, m, s); assert (m.value, s.value) == (1825044608, 1)
    lib.yfo_quantize_multiplier(alpha, m, s); assert (m.value, s.value) == (1460035712, -2)


def test_resolution_generalises(oracle):
    """The graph is fully convolutional (yolo_to_h5.py:134): 112x112 -> 14x14x18, 224x224 -> 28x28x18,
    and a 56x56 image embedded top-left in a larger canvas reproduces the interior cells it determines."""
    rng = np.random.default_rng(5)
    x = rng.integers(-128, 128, (224, 224, 3), dtype=np.int8)
    h224 = oracle.run(x)
    assert h224.shape == (28, 28, 18)
    assert oracle.run(x[:112, :112]).shape == (14, 14, 18)
    # top-left padding only => output cell (0,0)'s receptive field starts at the image origin
    assert np.array_equal(oracle.run(x[:56, :56])[0, 0], h224[0, 0])


def test_ragged_and_edge_inputs(oracle):
    for v in (-128, 127, 0):
        img = np.full((56, 56, 3), v, np.int8)
        h = oracle.run(img)
        assert h.shape == (7, 7, 18)
        # constant image on a big canvas: cells whose receptive field never touches a border agree
        big = oracle.run(np.full((224, 224, 3), v, np.int8))
        assert np.array_equal(big[12, 12], big[14, 15])
    out = oracle.run_batch(np.zeros((0, 56, 56, 3), np.int8), threads=4)
    assert out.shape == (0, 7, 7, 18)


def test_requant_primitives_vs_reference_cmsis_nn(oracle):
    """Row a11: the reference tree carries CMSIS-NN's statement of TFLite's two fixed-point primitives
    (stm32/Drivers/CMSIS/NN/Include/arm_nnsupportfunctions.h:210-263), compiled in place into oracle/_ref/libcmsis_ref.so
    (oracle/Makefile `ref`).  RoundingDivideByPOT is compared over the full sign range; the doubling-high-mult only for
    non-negative products: the header is ILP32 code and its `mult / (1UL << 31)` divides unsigned on an LP64 host."""
    import ctypes as C
    import os
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libcmsis_ref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libcmsis_ref.so not built (needs /root/reference at build time)")
    ref = C.CDLL(so)
    for f in (ref.ref_sat_doubling_high_mult, ref.ref_divide_by_power_of_two):
        f.restype = C.c_int32; f.argtypes = [C.c_int32, C.c_int32]
    rng = np.random.default_rng(11)
    xs = np.concatenate([rng.integers(-2**31, 2**31, 20000), np.arange(-4096, 4096), [2**31 - 1, -2**31, -2**31 + 1]])
    for e in (0, 1, 2, 5, 7, 8, 12, 20, 30, 31):
        for x in xs[:: 7 if e not in (1, 8) else 1]:
            assert oracle.lib.yfo_rdivpot(int(x), e) == ref.ref_divide_by_power_of_two(int(x), e), (x, e)
    mults = [1825044608, 1460035712, 1073741824, 2147483647, 1518500250] + rng.integers(2**30, 2**31, 40).tolist()
    accs = np.concatenate([rng.integers(0, 2**22, 4000), np.arange(0, 3000), rng.integers(0, 2**31, 2000)])
    n = 0
    for m in mults:
        for a in accs[::3]:
            assert oracle.lib.yfo_srdhm(int(a), int(m)) == ref.ref_sat_doubling_high_mult(int(a), int(m)), (a, m)
            n += 1
    # both negative -> non-negative product as well
    for a in rng.integers(-2**22, 0, 3000):
        m = -int(rng.integers(2**30, 2**31))
        assert oracle.lib.yfo_srdhm(int(a), m) == ref.ref_sat_doubling_high_mult(int(a), m)
    assert n > 50000
