"""float32 PyTorch restatement of the yoloface graph (BASELINE config 5).

Built from the int8 model itself: weights/biases are de-quantised (w*s_w[c], b*s_in*s_w[c]) and the
graph is walked op by op in float32 with torch ops.  Geometry follows the int8 graph / PyTorch
definition (yoloface/pytorch/yoloface.py:67-175: stride-2 convs pad top/left only) -- NOT the float
.tflite's SAME padding ("Trap").  Test infrastructure only."""
import numpy as np
import torch
import torch.nn.functional as F


def float_forward(oracle, img_int8):
    """img int8 [H,W,3] -> float32 logits [H/8,W/8,18]."""
    t_in = oracle.tensor(0)
    x = (torch.from_numpy(img_int8.astype(np.float32)) - t_in["zp"][0]) * np.float32(t_in["scale"][0])
    acts = {0: x.permute(2, 0, 1)[None]}                     # NCHW
    for i in range(oracle.num_ops):
        o = oracle.op(i); code = o["opcode"]; a = acts[o["inputs"][0]]
        if code == 34:      # PAD [[0,0],[t,b],[l,r],[0,0]] with real zero
            p = np.frombuffer(oracle.tensor(o["inputs"][1])["data"], "<i4")
            y = F.pad(a, (int(p[4]), int(p[5]), int(p[2]), int(p[3])))
        elif code in (3, 4):
            tw, tb = oracle.tensor(o["inputs"][1]), oracle.tensor(o["inputs"][2])
            s_in = np.float32(oracle.tensor(o["inputs"][0])["scale"][0])
            sw = np.array(tw["scale"], np.float32)
            w = np.frombuffer(tw["data"], np.int8).reshape(tw["shape"]).astype(np.float32)
            b = torch.from_numpy(np.frombuffer(tb["data"], "<i4").astype(np.float32) * s_in * sw)
            so = oracle.op_shape(i, img_int8.shape[0], img_int8.shape[1])[0]
            stride = 2 if so[1] * 2 <= a.shape[2] else 1
            if code == 3:   # OHWI -> OIHW
                wt = torch.from_numpy(w * sw[:, None, None, None]).permute(0, 3, 1, 2).contiguous()
                pad = (wt.shape[2] - 1) // 2 if (stride == 1 and wt.shape[2] > 1) else 0
                y = F.conv2d(a, wt, b, stride=stride, padding=pad)
            else:           # [1,KH,KW,C] -> [C,1,KH,KW]
                wt = torch.from_numpy(w * sw[None, None, None, :])[0].permute(2, 0, 1)[:, None].contiguous()
                y = F.conv2d(a, wt, b, stride=stride, padding=1 if stride == 1 else 0, groups=wt.shape[0])
        elif code == 98:
            y = F.leaky_relu(a, 0.10000000149011612)
        elif code == 17:
            so = oracle.op_shape(i, img_int8.shape[0], img_int8.shape[1])[0]
            k = 8 if a.shape[1] == 18 else 4                 # ops 8 and 25 (network.c:2648-2656, 2324-2332)
            y = F.max_pool2d(a, k, 2, padding=(k - 2) // 2)
            assert y.shape[2] == so[1]
        elif code == 0:
            y = a + acts[o["inputs"][1]]
        elif code == 114:
            y = a
        elif code == 2:
            y = torch.cat([acts[t] for t in o["inputs"]], dim=1)
        else:
            raise AssertionError(code)
        acts[o["output"]] = y
    return acts[100][0].permute(1, 2, 0).contiguous().numpy()


def decode_float(logits, conf_thr=0.7):
    """tflite_prediction.py:45-57 on float logits -> [n,5] (x1,y1,x2,y2,conf), candidate order as the firmware."""
    gh, gw, _ = logits.shape
    anchors = np.array([[9, 14], [12, 17], [22, 21]], np.float32)
    out = []
    sig = lambda v: 1.0 / (1.0 + np.exp(-v))
    for i in range(gh * gw):
        for j in range(3):
            t = logits.reshape(-1, 18)[i, j * 6:j * 6 + 6]
            conf = sig(t[4])
            if conf >= conf_thr:
                cx, cy = (sig(t[0]) + i % gw) * 8, (sig(t[1]) + i // gw) * 8
                w, h = np.exp(t[2]) * anchors[j, 0], np.exp(t[3]) * anchors[j, 1]
                out.append([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2, conf])
    return np.array(out, np.float32).reshape(-1, 5)
