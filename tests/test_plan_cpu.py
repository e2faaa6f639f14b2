"""CPU checks of the product's host logic: the C-ABI library loads and exports every declared
symbol, fails loudly without a GPU, regenerates ST's weight blob, and lowers the 54 TFLite ops to
26 fused steps whose tables reproduce the oracle bit-exactly (evaluated by tests/plan_emulator.py)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import pkg
from oracle_lib import vector_a, vector_b
from plan_emulator import Emulator

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def yf():
    m = pkg.load()
    m.build()
    return m


def test_library_exports_every_declared_symbol(yf):
    L = yf.lib()
    declared = set()
    for h in ("network.h", "network_data.h", "yoloface_b200.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        declared |= set(re.findall(r"AI_API_ENTRY\s+[\w\s\*]+?\b((?:ai_network|yf_b200)_\w+)\s*\(", src))
    assert len(declared) >= 27
    for name in declared:
        assert hasattr(L, name), name
    assert declared <= set(yf.EXPORTS)


def test_abi_struct_layout(yf):
    # ai_platform.h:517-525 (32 bytes on LP64), :467-470, :348-357
    assert C.sizeof(yf.AiBuffer) == 32 and yf.AiBuffer.data.offset == 16 and yf.AiBuffer.channels.offset == 12
    assert C.sizeof(yf.AiError) == 4 and C.sizeof(yf.AiNetworkParams) == 64
    assert yf.AI_BUFFER_FORMAT_S8 == 0x00840440 and yf.AI_BUFFER_FORMAT_U8 == 0x00040440


def test_no_cpu_fallback(yf):
    """Without a CUDA device the product must refuse to create a network (never compute on CPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(yf.AiRuntimeError) as ei:
        yf.Network()
    assert ei.value.type == 0x33      # AI_ERROR_CREATE_FAILED
    assert "no CPU path" in str(ei.value) or "CUDA" in str(ei.value)


def test_weights_blob_matches_st(yf, golden):
    """ai_network_data_weights_get() regenerates network_data.c's 11,304-byte blob from the .tflite."""
    blob = yf.weights_blob()
    assert blob == golden["st_blob"].tobytes()


def test_plan_structure(yf):
    P = yf.plan(56, 56)
    assert len(P["steps"]) == 26 and P["GH"] == 7 and P["GW"] == 7
    assert P["macs_per_image"] == 1029000
    kinds = [s["kind"] for s in P["steps"]]
    assert kinds.count(0) == 1 and kinds.count(1) == 16 and kinds.count(2) == 7 and kinds.count(3) == 2
    folded = sorted(o for s in P["steps"] for o in s["ops"])
    assert folded == [i for i in range(54) if i not in (22, 46)]   # concats are zero-copy
    assert P["n_epi"] == 544                                   # sum of conv/depthwise output channels
    # ST folds the same way: 31 c-nodes = 26 compute nodes + 3 eltwise + 2 concat (network.c:2193-2938)
    assert sum(1 for s in P["steps"] if s["add"][0]) == 3


@pytest.mark.parametrize("observer", [True, False])
def test_plan_tables_reproduce_oracle(yf, oracle, golden, observer):
    P = yf.plan(56, 56)
    emu = Emulator(P)
    imgs = [vector_a(), vector_b(), golden["images"][0], golden["images"][13],
            np.random.default_rng(3).integers(-128, 128, (56, 56, 3), dtype=np.int8)]
    for img in imgs:
        head, outs = oracle.run(img, dump=True)
        bufs = emu.run(img, observer=observer)
        assert np.array_equal(emu.tensor(bufs, 100), head)
        if observer:
            checked = 0
            for op in range(54):
                t = oracle.op(op)["output"]
                got = emu.tensor(bufs, t)
                if got is None or oracle.op(op)["opcode"] == 2:
                    continue
                assert np.array_equal(got, outs[op]), "op %d tensor %d" % (op, t)
                checked += 1
            assert checked == 49      # 54 ops - 3 PAD (folded) - 2 CONCAT (views)
            # concat outputs: slots hold the inputs in order
            for op in (22, 46):
                o = oracle.op(op)
                cat = np.concatenate([emu.tensor(bufs, t) for t in o["inputs"]], axis=-1)
                assert np.array_equal(cat, outs[op])


def test_plan_with_caller_supplied_blob(yf, oracle, golden):
    """Weights come from ai_network_init's params, not from the embedded model: perturb one weight
    in the ST blob and the plan must change accordingly."""
    blob = bytearray(golden["st_blob"].tobytes())
    P0 = yf.plan(56, 56, bytes(blob))
    assert np.array_equal(P0["wblob"], yf.plan(56, 56)["wblob"])
    blob[5] = (blob[5] + 1) & 0xFF
    P1 = yf.plan(56, 56, bytes(blob))
    assert not np.array_equal(P0["wblob"], P1["wblob"])


def test_plan_other_resolutions(yf, oracle):
    for H, W in ((112, 112), (224, 224), (64, 96)):
        P = yf.plan(H, W)
        assert (P["GH"], P["GW"]) == (H // 8, W // 8)
    P = yf.plan(112, 112)
    img = np.random.default_rng(9).integers(-128, 128, (112, 112, 3), dtype=np.int8)
    bufs = Emulator(P).run(img, observer=False)
    assert np.array_equal(Emulator(P).tensor(bufs, 100), oracle.run(img))
    with pytest.raises(RuntimeError):
        yf.plan(60, 56)


def test_fused_program_reproduces_oracle(yf, oracle, golden):
    """The single-kernel program (smem map by liveness, chunk-planar operands, per-phase parameter
    blocks) emulated on a flat, garbage-filled byte array gives the oracle's head bit-exactly."""
    from fused_emulator import run_fused
    F = yf.fused_program(56, 56)
    assert F["smem_bytes"] <= 113 * 1024          # two CTAs per SM for the generic kernel (descriptors in smem)
    assert F["smem_bytes_spec"] <= 75 * 1024      # three for the specialised one
    assert F["spec"] == 1                         # the compiled-in program IS this plan
    assert len(F["phases"]) == 26 and F["split"] == 14
    for seed, img in enumerate([vector_a(), vector_b(), golden["images"][5]]):
        head = run_fused(F, img, seed)
        assert np.array_equal(head.reshape(7, 7, 18), oracle.run(img))
    # image pairs: front phases twice, the twelve 7x7 phases once on the tall 16x7 image
    for seed, (i, j) in enumerate([(0, 1), (7, 3), (26, 26)]):
        a, b = golden["images"][i], golden["images"][j]
        ha, hb = run_fused(F, [a, b], 10 + seed)
        assert np.array_equal(ha.reshape(7, 7, 18), oracle.run(a)) and np.array_equal(hb.reshape(7, 7, 18), oracle.run(b))
    ha, hb = run_fused(F, [vector_a(), vector_b()], 5)
    assert np.array_equal(ha.reshape(7, 7, 18), oracle.run(vector_a())) and np.array_equal(hb.reshape(7, 7, 18), oracle.run(vector_b()))


def _has_rows(warp, t0, nt, rows, chunks, wgs):
    """yf_plan.h fused_has_rows(), restated."""
    wg, q = warp >> 2, warp & 3
    if nt >= wgs:
        return (t0 + wg) * 128 + q * 32 < rows
    return any((t0 + u // chunks) * 128 + q * 32 < rows for u in range(wg, nt * chunks, wgs))


@pytest.mark.parametrize("threads", [256, 512, 4512])
def test_fused_program_cta_shapes(yf, oracle, golden, threads):
    """The CTA shapes of the fused kernel (256 threads: throughput; 512 threads, all of TMEM: latency; 4512: latency with
    a cluster of 4 CTAs sharing an image's front phases) come from the same planner: the emulated program gives the
    oracle's head, every (tile, chunk) unit of every conv phase has exactly one owning warpgroup (of one CTA), and the
    barrier counts equal the warps that own rows plus the control warp."""
    from fused_emulator import run_fused
    F = yf.fused_program(56, 56, threads=threads)
    cluster, nthreads = (threads // 1000 or 1), threads % 1000
    wgs = nthreads // 128
    assert F["threads"] == nthreads and F["cluster"] == cluster and F["warpgroups"] == wgs and F["tmem_cols"] == (128 if nthreads == 256 else 512)
    assert F["spec"] == 1                          # a specialised kernel exists for each shape
    assert len(F["phases"]) == 26 and F["split"] == 14
    img = golden["images"][11]
    assert np.array_equal(run_fused(F, img, 3).reshape(7, 7, 18), oracle.run(img))
    ctrl = 4 * wgs - 1
    for pi, ph in enumerate(F["phases"]):
        shared = cluster > 1 and pi < F["split"]    # front phases of the cluster shape are dealt over the whole cluster
        if ph["kind"] == 2:
            assert ph_per(ph) == (cluster if shared else 1) * nthreads // ph["nw"]
        if ph["kind"] != 1:
            continue
        if shared:
            assert ph["ntiles"] <= ph["tpg"]        # one tile group; masks / counts are indexed by the CTA's rank
            nt, chunks, rows, W = ph["ntiles"], ph["chunks_out"], ph["rows_out"], wgs * cluster
            seen = {}
            for r in range(cluster):
                owners = [w for w in range(4 * wgs) if _has_rows((((w >> 2) * cluster + r) << 2) | (w & 3), 0, nt, rows, chunks, W)]
                mask = (ph["own" + str(r >> 1)] >> (16 * (r & 1))) & 0xffff
                assert mask == sum(1 << w for w in owners)
                assert ((ph["grp_warps"] >> (8 * r)) & 0xff) == len([w for w in owners if w != ctrl]) + 1
                for w in owners:
                    seen[(r, w >> 2)] = True
            for t in range(nt):                     # every unit with real rows: exactly one (rank, warpgroup)
                for q in range(4):
                    if t * 128 + q * 32 >= rows:
                        continue
                    for c in range(chunks):
                        own = [v for v in range(W) if (t in range(v, nt, W) if nt >= W else (t * chunks + c) in range(v, nt * chunks, W))]
                        assert len(own) == 1 and (own[0] % cluster, own[0] // cluster) in seen
            continue
        for rows, counts, key in ((ph["rows_out"], ph["grp_warps"], "own"), (ph["rows_single"], ph["grp_warps_single"], "own_single")):
            for g, t0 in enumerate(range(0, ph["ntiles"], ph["tpg"])):
                nt = min(ph["tpg"], ph["ntiles"] - t0)
                owners = [w for w in range(4 * wgs) if _has_rows(w, t0, nt, rows, ph["chunks_out"], wgs)]
                assert ((counts >> (8 * g)) & 0xff) == len(owners) + (0 if ctrl in owners else 1)
                mask = (ph[key + str(g >> 1)] >> (16 * (g & 1))) & 0xffff
                assert mask == sum(1 << w for w in owners)        # the kernel tests these bits instead of re-deriving them
                # every unit with real rows is reached by exactly one warpgroup (per lane quarter)
                for t in range(nt):
                    for q in range(4):
                        if (t0 + t) * 128 + q * 32 >= rows:
                            continue
                        for c in range(ph["chunks_out"]):
                            if nt >= wgs:
                                own = [wg for wg in range(wgs) if t in range(wg, nt, wgs)]
                            else:
                                own = [wg for wg in range(wgs) if (t * ph["chunks_out"] + c) in range(wg, nt * ph["chunks_out"], wgs)]
                            assert len(own) == 1 and (4 * own[0] + q) in owners


def test_fused_program_rejects_unknown_shapes(yf):
    """CTAs of 256 or 512 threads; clusters (of 2 or 4) only of 512-thread CTAs.  A cluster of 2 can be laid out but has
    no compiled kernel (spec == 0)."""
    for threads in (128, 384, 1024, 8512, 4256, 3512):
        with pytest.raises(RuntimeError):
            yf.fused_program(56, 56, threads=threads)
    assert yf.fused_program(56, 56, threads=2512)["spec"] == 0
    # the cluster shape needs every front conv layer's tiles in ONE TMEM group: true for the deployed model, and the
    # layout is the plain latency shape's apart from the dealing fields
    a, b = yf.fused_program(56, 56, threads=512), yf.fused_program(56, 56, threads=4512)
    assert a["arena_bytes"] == b["arena_bytes"] and a["smem_bytes_spec"] == b["smem_bytes_spec"]
    assert np.array_equal(a["params"], b["params"])
    for pa, pb in zip(a["phases"], b["phases"]):
        for k in pa:
            if k not in ("per", "grp_warps", "grp_warps_single", "own0", "own1", "own_single0", "own_single1"):
                assert pa[k] == pb[k], k


def ph_per(ph):
    return ph["per"]


@pytest.mark.parametrize("hw", [(8, 8), (8, 16), (24, 16), (32, 32), (40, 56), (56, 64), (64, 64), (16, 128)])
def test_fused_program_other_resolutions(yf, oracle, hw):
    """The allocator, the word-plane skew, the tile groups and the mul-shift divisions all depend on the shape:
    tiny inputs (pool windows larger than the tensor, single-tile layers) up to the largest fused size."""
    from fused_emulator import run_fused
    H, W = hw
    F = yf.fused_program(H, W)
    rng = np.random.default_rng(H * 1000 + W)
    img, img2 = rng.integers(-128, 128, (2, H, W, 3), dtype=np.int8)
    assert np.array_equal(run_fused(F, img, 1).reshape(H // 8, W // 8, 18), oracle.run(img))
    if F["split"] < len(F["phases"]):
        ha, hb = run_fused(F, [img, img2], 2)
        assert np.array_equal(ha.reshape(H // 8, W // 8, 18), oracle.run(img)) and np.array_equal(hb.reshape(H // 8, W // 8, 18), oracle.run(img2))


def test_fused_program_limits(yf):
    F = yf.fused_program(64, 64)                  # still fits shared memory and TMEM
    assert F["smem_bytes"] < 200 * 1024
    for hw in ((64, 96), (224, 224)):             # a layer's accumulator tiles exceed TMEM / smem -> layered path
        with pytest.raises(RuntimeError):
            yf.fused_program(*hw)


def test_st_activation_mode_reproduces_st_tables(yf, oracle, golden, monkeypatch):
    """with ST-style activations the 17 LeakyReLU tables equal network.c:2218..2902 exactly;
    the default (TFLite) tables differ from them in 271 entries."""
    def leaky_tables(P):
        out = {}
        for s in P["steps"]:
            for op in s["ops"]:
                if oracle.op(op)["opcode"] == 98:
                    out[op] = P["luts"][s["lut1"]]
        return out
    st = dict(zip(golden["st_lut_ops"].tolist(), golden["st_luts"]))
    tfl = leaky_tables(yf.plan(56, 56))
    assert sum(int((tfl[op] != st[op]).sum()) for op in st) == 271
    monkeypatch.setenv("YF_B200_ST_ACTIVATIONS", "1")
    stm = leaky_tables(yf.plan(56, 56))
    assert sorted(stm) == sorted(st)
    for op in st:
        assert np.array_equal(stm[op], st[op]), op
    # the three QUANTIZE operators ST folds into forward_concat (network.c:2307-2313, 2631-2637): the ST mode builds their
    # tables with the same float32 / round-half-even rule as ST's activation tables (a reading, not a pin: the concat's
    # own arithmetic is inside the closed library)
    def quant_tables(P):
        out = {}
        for s in P["steps"]:
            qs = [op for op in s["ops"] if oracle.op(op)["opcode"] == 114]
            if qs:
                out[qs[0]] = P["luts"][s["lut2"] if s["lut2"] >= 0 else s["lut1"]]
        return out
    qst = quant_tables(yf.plan(56, 56))
    assert sorted(qst) == [21, 44, 45]
    for op, tab in qst.items():
        o = oracle.op(op); ti, to = oracle.tensor(o["inputs"][0]), oracle.tensor(o["output"])
        q = np.arange(-128, 128, dtype=np.float32)
        v = ((q - np.float32(ti["zp"][0])) * np.float32(ti["scale"][0])) / np.float32(to["scale"][0])
        want = np.clip(np.rint(v).astype(np.int64) + to["zp"][0], -128, 127).astype(np.int8)
        assert np.array_equal(tab.view(np.int8), want), op
    # and the plan still executes: heads move by a few LSB at most relative to TFLite mode
    img = golden["images"][2]
    a = Emulator(yf.plan(56, 56)).run(img, observer=False)
    monkeypatch.delenv("YF_B200_ST_ACTIVATIONS")
    b = Emulator(yf.plan(56, 56)).run(img, observer=False)
    d = np.abs(a[1].astype(int) - b[1].astype(int))
    assert 0 < d.max() <= 16
