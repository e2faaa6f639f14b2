"""numpy interpreter of the product's *plan* (yf_b200_plan_json / yf_b200_plan_blob).

It evaluates each fused step with exactly the data layout and folded arithmetic the CUDA kernels
use (channel-pitched buffers, concat slots, permuted/padded weight images, EpiCh requantisation,
256-entry tables, fused ADD), so that the lowering in csrc/yf_plan.cc can be checked against the
CPU oracle without a GPU.  Test infrastructure only."""
import numpy as np


def requant(acc, epi):
    """acc int64 [..., C]; epi structured array [C] -> int64 (not clamped, includes zp_out)."""
    p = (acc << epi["ls"].astype(np.int64)) * epi["mult"].astype(np.int64) + epi["add64"]
    t = p >> 31
    s = (t >> 31) & epi["sgn_mask"].astype(np.int64)
    return (t + epi["c2"].astype(np.int64) + s) >> epi["e"].astype(np.int64)


def mbqm(x, m, s):
    x = x.astype(np.int64)
    t = (x * int(m) + (1 << 30)) >> 31
    rs = -int(s)
    if rs == 0:
        return t
    return (t + (1 << (rs - 1)) + (t >> 31)) >> rs


def lut_apply(v, table):
    return table[v.astype(np.int64) + 128].astype(np.int64)


class Emulator:
    def __init__(self, plan):
        self.P = plan

    def run(self, img, observer=True):
        """img int8 [H,W,3] -> dict buffer index -> int8 array [H,W,CP]."""
        P = self.P
        bufs = {}
        for i, b in enumerate(P["buffers"]):
            if b["is_input"]:
                bufs[i] = img.astype(np.int8)
            else:
                bufs[i] = np.zeros((b["H"], b["W"], b["CP"]), np.int8)
        wblob = P["wblob"]
        for s in P["steps"]:
            kind = s["kind"]
            x = bufs[s["in_buf"]]
            cout, npad, kpad = s["Cout"], s["Npad"], s["Kpad"]
            epi = P["epi"][s["epi_base"]:s["epi_base"] + cout] if s["epi_base"] >= 0 else None
            if kind == 1:  # conv1x1: raw int8 activations (pitch CP) x packed weights [K/16][Npad][16]
                cp = x.shape[2]
                wimg = wblob[s["w_off"]:s["w_off"] + s["w_bytes"]].view(np.int8).reshape(kpad // 16, npad, 16)
                wmat = wimg.transpose(0, 2, 1).reshape(kpad, npad).astype(np.int64)       # [K, N]
                a = np.zeros((x.shape[0], x.shape[1], kpad), np.int64); a[..., :cp] = x
                acc = a @ wmat
                y = np.clip(requant(acc[..., :cout], epi), -128, 127)
            elif kind == 0:  # im2col conv 3x3 s2, pad top/left with in_zp
                H, W, _ = x.shape
                xp = np.full((H + 1, W + 1, 3), s["in_zp"], np.int64); xp[1:, 1:] = x
                wimg = wblob[s["w_off"]:s["w_off"] + s["w_bytes"]].view(np.int8).reshape(kpad // 16, npad, 16)
                wmat = wimg.transpose(0, 2, 1).reshape(kpad, npad).astype(np.int64)
                Ho, Wo = s["Hout"], s["Wout"]
                a = np.zeros((Ho, Wo, kpad), np.int64)
                for ky in range(3):
                    for kx in range(3):
                        a[..., (ky * 3 + kx) * 3:(ky * 3 + kx) * 3 + 3] = xp[ky:ky + 2 * Ho:2, kx:kx + 2 * Wo:2]
                acc = a @ wmat
                y = np.clip(requant(acc[..., :cout], epi), -128, 127)
            elif kind == 2:  # depthwise 3x3: one-hot dp4a words [9][CP]
                cp = x.shape[2]
                w1h = wblob[s["w_off"]:s["w_off"] + s["w_bytes"]].view(np.uint32).reshape(9, cp)
                w = np.zeros((9, cp), np.int64)
                for c in range(cp):
                    w[:, c] = ((w1h[:, c] >> (8 * (c % 4))) & 0xFF).astype(np.uint8).view(np.int8)
                    assert np.all((w1h[:, c] & ~np.uint32(0xFF << (8 * (c % 4)))) == 0)
                H, W, _ = x.shape
                Ho, Wo, st = s["Hout"], s["Wout"], s["stride"]
                xp = np.full((H + 2 + st, W + 2 + st, cp), s["in_zp"], np.int64)
                xp[s["pad_t"]:s["pad_t"] + H, s["pad_l"]:s["pad_l"] + W] = x
                acc = np.zeros((Ho, Wo, cp), np.int64)
                for ky in range(3):
                    for kx in range(3):
                        acc += xp[ky:ky + st * Ho:st, kx:kx + st * Wo:st] * w[ky * 3 + kx]
                y = np.clip(requant(acc[..., :cout], epi), -128, 127)
            elif kind == 3:  # maxpool over valid cells
                H, W, _ = x.shape
                Ho, Wo, st, k = s["Hout"], s["Wout"], s["stride"], s["kh"]
                y = np.full((Ho, Wo, cout), -128, np.int64)
                for oy in range(Ho):
                    for ox in range(Wo):
                        y0, x0 = oy * st - s["pad_t"], ox * st - s["pad_l"]
                        win = x[max(0, y0):min(H, y0 + k), max(0, x0):min(W, x0 + k), :cout]
                        y[oy, ox] = win.reshape(-1, cout).max(axis=0)
            elif kind == 4:
                y = x[..., s["in_coff"]:s["in_coff"] + cout].astype(np.int64)
            else:
                raise AssertionError(kind)
            raw = y
            add = s["add"]
            if add[0]:
                _, zp1, zp2, zpo, m1, m2, mo, s1, s2, so = add
                other = bufs[s["add_buf"]][..., s["add_coff"]:s["add_coff"] + cout].astype(np.int64)
                if observer and s["pre_add_buf"] >= 0:
                    bufs[s["pre_add_buf"]][..., :cout] = raw
                sx = mbqm((other - zp1) << 20, m1, s1); sy = mbqm((y - zp2) << 20, m2, s2)
                y = np.clip(mbqm(sx + sy, mo, so) + zpo, -128, 127)
            elif observer and s["raw_buf"] >= 0:
                bufs[s["raw_buf"]][..., :cout] = raw
            if observer:
                if s["lut1"] >= 0:
                    y = lut_apply(y, P["luts"][s["lut1"]])
                if s["mid_buf"] >= 0:
                    bufs[s["mid_buf"]][..., :cout] = y
                if s["lut2"] >= 0:
                    y = lut_apply(y, P["luts"][s["lut2"]])
            elif s["lut_fused"] >= 0:
                y = lut_apply(y, P["luts"][s["lut_fused"]])
            bufs[s["out_buf"]][..., s["out_coff"]:s["out_coff"] + cout] = y
        return bufs

    def tensor(self, bufs, t):
        buf, coff, C = self.P["loc"][t]
        if buf < 0:
            return None
        return bufs[buf][..., coff:coff + C]
