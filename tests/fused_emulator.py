"""numpy emulation of the fused single-kernel program (yf_b200_fused_json + parameter blob).

Shared memory is one flat byte array pre-filled with random garbage and addressed with the very
offsets / strides the kernel uses (UMMA operand rows at in_off + chunk*in_cs + row*16 -- including the
rows and K chunks an MMA tile reads past its buffer, which must meet zero weights --, word-planar
bordered buffers at off + word*ws + cell*4, weight images [K/16][N][16] in the parameter slot,
aliasing by liveness), so that layout, aliasing or parameter-block mistakes in
csrc/yf_plan.cc::build_fused surface on CPU.
Test infrastructure only."""
import numpy as np

from plan_emulator import lut_apply, mbqm, requant


def lean_to_epi(bias9, mult, kc, sh, dtype):
    """the kernels' 16-byte requant constants (csrc/yf_requant.cuh: bias' << 9, m, 2^7 + 256 * c2p, 8 + e) -> EpiCh records"""
    n = len(mult)
    epi = np.zeros(n, dtype)
    assert np.all(bias9 % 512 == 0) and np.all((kc - 128) % 256 == 0) and np.all(sh >= 9)
    epi["mult"] = mult
    epi["add64"] = (bias9.astype(np.int64) // 512) * mult.astype(np.int64) + (1 << 30)
    epi["e"] = sh - 8
    epi["c2"] = (kc.astype(np.int64) - 128) // 256 - (128 << epi["e"].astype(np.int64))
    epi["sgn_mask"] = -1
    return epi


def run_fused(F, img, seed=0):
    """One image (a pair holding only image A) or a list of two images (front phases twice, back phases once on the
    tall pair image).  Returns the head, or the list of two heads."""
    imgs = img if isinstance(img, (list, tuple)) else [img]
    assert 1 <= len(imgs) <= 2
    rng = np.random.default_rng(seed)
    smem = rng.integers(0, 256, F["smem_bytes"] + (1 << 16), dtype=np.uint8)
    params, epi_all = F["params"], F["epi"]
    split = F.get("split", len(F["phases"]))
    heads = [None, None]
    pair_b = len(imgs) == 2
    if split == len(F["phases"]):
        assert len(imgs) == 1, "this program has no pair phases"
    for k, im in enumerate(imgs):
        smem[F["in_off"]:F["in_off"] + F["in_bytes"]] = np.ascontiguousarray(im).view(np.uint8).reshape(-1)
        for ph in F["phases"][:split]:
            _phase(F, ph, smem, rng, params, epi_all, heads, shift=ph.get("out_pair_shift", 0) if k else 0, pair_b=False)
        # the front phases of the next image may scribble over everything but the buffers that cross into the back
    for ph in F["phases"][split:]:
        _phase(F, ph, smem, rng, params, epi_all, heads, shift=0, pair_b=pair_b)
    return heads if pair_b else heads[0]


def _phase(F, ph, smem, rng, params, epi_all, heads, shift, pair_b):

    def chunk_rows(off, cs, chunk, rows):          # -> int8 [rows,16] view of one chunk
        a = off + chunk * cs
        return smem[a:a + rows * 16].view(np.int8).reshape(rows, 16)

    def word_plane(off, ws, word, cells):          # -> int8 [cells,4] view of one word plane
        a = off + word * ws
        return smem[a:a + cells * 4].view(np.int8).reshape(cells, 4)

    def padded_input(ph, chunks):                  # word-planar zero-point-bordered (H+2)x(W+2) buffer -> [H+2, W+2, nw*4] int64
        H, W, nw = ph["Hin"], ph["Win"], ph["nw"]
        cells = (H + 2) * (W + 2)
        assert ph["in_wp"] == W + 2 and ph["in_ws"] >= cells * 4 and ph["in_ws"] % 4 == 0
        return np.concatenate([word_plane(ph["in_off"], ph["in_ws"], w, cells) for w in range(nw)], axis=1).reshape(H + 2, W + 2, nw * 4).astype(np.int64)

    if True:
        slot = params[ph["param_off"]:ph["param_off"] + ph["param_bytes"]]
        kind, cout, npad = ph["kind"], ph["cout"], ph["npad"]
        pair = ph.get("pair", 0)
        rows_o = ph["rows_out"] if (not pair or pair_b) else ph["rows_single"]     # rows the conv epilogues process
        rows_dw = ph["rows_out"] if (not pair or pair_b) else ph["rows_a"]         # rows the depthwise phases produce
        lut = slot[ph["lut_off"]:ph["lut_off"] + 256].view(np.int8) if ph["has_lut"] else None
        if ph["scratch_off"] >= 0:      # the kernel scribbles here during this phase: must not alias anything live
            size = 2 * F["warpgroups"] * 6144 + 2048 if kind == 0 else ph["nw"] * ph["scratch_ws"]
            assert kind == 0 or ph["scratch_ws"] >= ph["Hin"] * ph["Wout"] * 4
            smem[ph["scratch_off"]:ph["scratch_off"] + size] = rng.integers(0, 256, size, dtype=np.uint8)
        if kind in (0, 1):
            K = ph["nk"] * 32
            wimg = slot[ph["w_off"]:ph["w_off"] + (K // 16) * npad * 16].view(np.int8).reshape(K // 16, npad, 16)
            wmat = wimg.transpose(0, 2, 1).reshape(K, npad).astype(np.int64)
            if kind == 1:
                tile_rows = ph["ntiles"] * 128                  # what the MMAs really read: whole 128-row tiles of every K chunk
                assert ph["in_off"] + (K // 16 - 1) * ph["in_cs"] + tile_rows * 16 <= F["smem_bytes"], "MMA operand read leaves the allocation"
                assert 1 <= ph["tpg"] and ph["tpg"] * npad <= F["tmem_cols"]
                A = np.concatenate([chunk_rows(ph["in_off"], ph["in_cs"], c, tile_rows) for c in range(K // 16)], axis=1).astype(np.int64)[:rows_o]
            else:
                H, W = ph["Hin"], ph["Win"]
                x = smem[F["in_off"]:F["in_off"] + H * W * 3].view(np.int8).reshape(H, W, 3).astype(np.int64)
                xp = np.full((H + 1, W + 3, 3), ph["in_zp"], np.int64); xp[1:, 1:W + 1] = x
                flat = xp.reshape(H + 1, -1)
                A = rng.integers(-128, 128, (rows_o, 64)).astype(np.int64)       # don't-care bytes are garbage
                Ho, Wo = ph["Hout"], ph["Wout"]
                for r in range(rows_o):
                    oy, ox = divmod(r, Wo)
                    for ky in range(3):
                        A[r, ky * 16:ky * 16 + 9] = flat[2 * oy + ky, (2 * ox) * 3:(2 * ox) * 3 + 9]
            acc = A @ wmat
            raw = slot[ph["epi_off"]:ph["epi_off"] + cout * 16].reshape(cout, 16)      # constants travel in the block: {bias' << 9, mult, kc, 8 + e}
            epi = lean_to_epi(raw[:, 0:4].copy().view("<i4").reshape(-1), raw[:, 4:8].copy().view("<i4").reshape(-1),
                              raw[:, 8:12].copy().view("<i4").reshape(-1), raw[:, 12:16].copy().view("<i4").reshape(-1), epi_all.dtype)
            ref = epi_all[ph["epi_base"]:ph["epi_base"] + cout]
            assert all(np.array_equal(epi[f], ref[f]) for f in ("add64", "mult", "e", "c2"))
            y = np.clip(requant(acc[:, :cout], epi), -128, 127)
            if ph["add_off"] >= 0:
                _, zp1, zp2, zpo, m1, m2, mo, s1, s2, so = ph["add"]
                skip = np.concatenate([chunk_rows(ph["add_off"], ph["add_cs"], g, rows_o) for g in range(ph["chunks_out"])], axis=1)[:, :cout].astype(np.int64)
                y = np.clip(mbqm(mbqm((skip - zp1) << 20, m1, s1) + mbqm((y - zp2) << 20, m2, s2), mo, so) + zpo, -128, 127)
            elif lut is not None:
                y = lut_apply(y, lut)
        elif kind == 2:
            chunks = ph["chunks_out"]; cp = chunks * 16
            w1h = slot[ph["dw_off"]:ph["dw_off"] + 9 * cp * 4].view(np.uint32).reshape(9, cp)
            w = np.zeros((9, cp), np.int64)
            for c in range(cp):
                w[:, c] = ((w1h[:, c] >> (8 * (c % 4))) & 0xFF).astype(np.uint8).view(np.int8)
            nch = ph["nw"] * 4
            raw = slot[ph["dwepi_off"]:ph["dwepi_off"] + nch * 16].copy().view("<i4").reshape(4, nch)   # [bias9 | mult | kc | sh][nw*4]
            epi = lean_to_epi(raw[0][:cout], raw[1][:cout], raw[2][:cout], raw[3][:cout], epi_all.dtype)
            ref = epi_all[ph["epi_base"]:ph["epi_base"] + cout]
            assert all(np.array_equal(epi[f], ref[f]) for f in ("add64", "mult", "e", "c2"))
            H, W, Ho, Wo, st = ph["Hin"], ph["Win"], ph["Hout"], ph["Wout"], ph["stride"]
            xp = padded_input(ph, chunks)           # the border must already hold the zero point (written by the producer)
            w = w[:, :xp.shape[2]]
            assert np.all(xp[0, :, :cout] == ph["in_zp"]) and np.all(xp[:, 0, :cout] == ph["in_zp"]) and np.all(xp[-1, :, :cout] == ph["in_zp"]) and np.all(xp[:, -1, :cout] == ph["in_zp"])
            if pair:                                # the separator rows are image A's bottom / image B's top border
                assert np.all(xp[ph["sep_y"] + 1, :, :cout] == ph["in_zp"])
                assert not pair_b or np.all(xp[ph["sep_y"] + 2, :, :cout] == ph["in_zp"])
            oy0, ox0 = 1 - ph["pad_t"], 1 - ph["pad_l"]
            acc = np.zeros((Ho, Wo, xp.shape[2]), np.int64)
            for ky in range(3):
                for kx in range(3):
                    acc += xp[oy0 + ky:oy0 + ky + st * Ho:st, ox0 + kx:ox0 + kx + st * Wo:st] * w[ky * 3 + kx]
            y = np.clip(requant(acc.reshape(Ho * Wo, -1)[:, :cout], epi), -128, 127)
            if lut is not None:
                y = lut_apply(y, lut)
            y = y[:rows_dw]
        elif kind == 3:
            chunks = ph["chunks_out"]; cp = chunks * 16
            H, W, Ho, Wo, st, k = ph["Hin"], ph["Win"], ph["Hout"], ph["Wout"], ph["stride"], ph["ksize"]
            x = padded_input(ph, chunks)[1:-1, 1:-1]
            y = np.zeros((Ho * Wo, cout), np.int64)
            for oy in range(Ho):
                for ox in range(Wo):
                    y0, x0 = oy * st - ph["pad_t"], ox * st - ph["pad_l"]
                    y[oy * Wo + ox] = x[max(0, y0):min(H, y0 + k), max(0, x0):min(W, x0 + k), :cout].reshape(-1, cout).max(axis=0)
            if lut is not None:
                y = lut_apply(y, lut)
        else:
            raise AssertionError(kind)
        nrows = y.shape[0]
        if ph["to_global"]:
            if pair:
                heads[0] = y[:ph["rows_a"]].astype(np.int8)
                if pair_b:
                    heads[1] = y[ph["row_b0"]:ph["row_b0"] + ph["rows_a"]].astype(np.int8)
            else:
                heads[0] = y.astype(np.int8)
            return
        out = rng.integers(-128, 128, (nrows, ph["chunks_out"] * 16)).astype(np.int8)   # pad channels: garbage (must meet zero weights)
        out[:, :cout] = y
        if ph["out_wp"]:                            # producer writes the interior of a word-planar bordered buffer and fills the border
            Ho, Wo, nw = ph["Hout"], ph["Wout"], ph["nw"]
            cells = (Ho + 2) * (Wo + 2)
            assert ph["out_ws"] >= cells * 4 and ph["out_wp"] == Wo + 2 and shift == 0
            for w in range(nw):                     # start from what is there: rows the kernel does not process keep their bytes
                a = ph["out_off"] + w * ph["out_ws"]
                plane = smem[a:a + cells * 4].view(np.int8).reshape(Ho + 2, Wo + 2, 4)
                plane[0] = ph["out_zp"]; plane[-1] = ph["out_zp"]; plane[:, 0] = ph["out_zp"]; plane[:, -1] = ph["out_zp"]
                for r in range(nrows):
                    yy, xx = divmod(r, Wo)
                    sep = pair and yy in (ph["sep_y"], ph["sep_y"] + 1)
                    plane[yy + 1, xx + 1] = ph["out_zp"] if sep else out[r, w * 4:(w + 1) * 4]
            return
        for g in range(ph["chunks_out"]):
            a = ph["out_off"] + shift + g * ph["out_cs"]
            smem[a:a + nrows * 16] = out[:, g * 16:(g + 1) * 16].reshape(-1).view(np.uint8)
