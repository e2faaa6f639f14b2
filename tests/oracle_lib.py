"""ctypes binding of the CPU oracle (oracle/liboracle.so).  Test infrastructure only:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
MODEL_PATH = os.path.join(ROOT, "stm32h7-yolo_b200", "assets", "yoloface_int8.tflite")


class Det(C.Structure):
    _fields_ = [("x1", C.c_float), ("y1", C.c_float), ("x2", C.c_float), ("y2", C.c_float), ("conf", C.c_float)]


def build_oracle():
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    src = os.path.join(ORACLE_DIR, "yf_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])
    return so


class Oracle:
    def __init__(self, model_path=MODEL_PATH):
        self.lib = lib = C.CDLL(build_oracle())
        self.buf = open(model_path, "rb").read()
        self._cbuf = C.create_string_buffer(self.buf, len(self.buf))
        lib.yfo_load.restype = C.c_void_p
        lib.yfo_load.argtypes = [C.c_void_p, C.c_size_t]
        lib.yfo_last_error.restype = C.c_char_p
        lib.yfo_free.argtypes = [C.c_void_p]
        for f in ("yfo_num_tensors", "yfo_num_ops", "yfo_input_tensor", "yfo_output_tensor"):
            getattr(lib, f).argtypes = [C.c_void_p]
        lib.yfo_tensor_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                        C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        lib.yfo_tensor_scale.restype = C.c_float
        lib.yfo_tensor_scale.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.yfo_tensor_zp.restype = C.c_int64
        lib.yfo_tensor_zp.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.yfo_tensor_name.restype = C.c_char_p
        lib.yfo_tensor_name.argtypes = [C.c_void_p, C.c_int]
        lib.yfo_op_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.yfo_op_out_elems.restype = C.c_long
        lib.yfo_op_out_elems.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
        lib.yfo_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
        lib.yfo_run_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        lib.yfo_quantize_multiplier.argtypes = [C.c_double, C.POINTER(C.c_int32), C.POINTER(C.c_int)]
        for f in ("yfo_srdhm",):
            getattr(lib, f).restype = C.c_int32
            getattr(lib, f).argtypes = [C.c_int32, C.c_int32]
        lib.yfo_rdivpot.restype = C.c_int32
        lib.yfo_rdivpot.argtypes = [C.c_int32, C.c_int]
        lib.yfo_mbqm.restype = C.c_int32
        lib.yfo_mbqm.argtypes = [C.c_int32, C.c_int32, C.c_int]
        lib.yfo_leaky_lut.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.yfo_decode_nms.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float, C.c_float, C.c_int,
                                       C.POINTER(Det), C.c_int]
        lib.yfo_decode_nms_ex.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_float, C.c_float,
                                          C.c_float, C.c_int, C.POINTER(Det), C.c_int]
        lib.yfo_decode_all.restype = None
        lib.yfo_decode_all.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_float, C.POINTER(Det)]
        lib.yfo_rgb565_to_input.argtypes = [C.c_void_p, C.c_void_p]
        self.m = lib.yfo_load(self._cbuf, len(self.buf))
        if not self.m:
            raise RuntimeError("oracle load failed: %s" % lib.yfo_last_error().decode())
        self.num_ops = lib.yfo_num_ops(self.m)
        self.num_tensors = lib.yfo_num_tensors(self.m)

    def __del__(self):
        try:
            self.lib.yfo_free(self.m)
        except Exception:
            pass

    # ---- model introspection -------------------------------------------------------------
    def tensor(self, t):
        shape = (C.c_int * 4)(); ty = C.c_int(); ns = C.c_int(); qd = C.c_int()
        data = C.c_void_p(); dl = C.c_size_t()
        rank = self.lib.yfo_tensor_info(self.m, t, shape, ty, ns, qd, C.byref(data), C.byref(dl))
        scales = [self.lib.yfo_tensor_scale(self.m, t, i) for i in range(ns.value)]
        zps = [self.lib.yfo_tensor_zp(self.m, t, i) for i in range(ns.value)]
        raw = C.string_at(data.value, dl.value) if data.value else b""
        return dict(shape=list(shape)[:rank], type=ty.value, scale=scales, zp=zps, qdim=qd.value, data=raw,
                    name=self.lib.yfo_tensor_name(self.m, t).decode())

    def op(self, i):
        code = C.c_int(); ins = (C.c_int * 3)(); out = C.c_int()
        n = self.lib.yfo_op_info(self.m, i, code, ins, out)
        return dict(opcode=code.value, inputs=list(ins)[:n], output=out.value)

    def op_shape(self, i, H=56, W=56):
        sh = (C.c_int * 4)()
        n = self.lib.yfo_op_out_elems(self.m, i, H, W, sh)
        return list(sh), n

    # ---- execution ------------------------------------------------------------------------
    def run(self, img, dump=False):
        """img: int8 [H,W,3] -> head int8 [H/8,W/8,18] (and, with dump=True, a list of every op's output)."""
        img = np.ascontiguousarray(img, dtype=np.int8)
        H, W, _ = img.shape
        out = np.empty((H // 8, W // 8, 18), np.int8)
        outs = None; ptrs = None
        if dump:
            outs = []
            ptrs = (C.c_void_p * self.num_ops)()
            for i in range(self.num_ops):
                sh, n = self.op_shape(i, H, W)
                a = np.empty(sh[1:], np.int8); outs.append(a); ptrs[i] = a.ctypes.data
        rc = self.lib.yfo_run(self.m, img.ctypes.data, H, W, out.ctypes.data, ptrs)
        if rc:
            raise RuntimeError(self.lib.yfo_last_error().decode())
        return (out, outs) if dump else out

    def run_batch(self, imgs, threads=1):
        imgs = np.ascontiguousarray(imgs, dtype=np.int8)
        n, H, W, _ = imgs.shape
        out = np.empty((n, H // 8, W // 8, 18), np.int8)
        rc = self.lib.yfo_run_batch(self.m, imgs.ctypes.data, n, H, W, out.ctypes.data, threads)
        if rc:
            raise RuntimeError(self.lib.yfo_last_error().decode())
        return out

    def leaky_lut(self, op):
        lut = np.empty(256, np.int8)
        if self.lib.yfo_leaky_lut(self.m, op, lut.ctypes.data):
            raise ValueError("op %d is not LEAKY_RELU" % op)
        return lut

    def decode_nms(self, head, conf_thr=0.7, iou_thr=0.4, plus_one=False, scale=0.14218327403068542, zp=-15,
                   anchors=None, stride=8.0, max_det=None):
        head = np.ascontiguousarray(head, dtype=np.int8)
        gh, gw, _ = head.shape
        cap = gh * gw * 3 if max_det is None else max_det
        dets = (Det * max(cap, 1))()
        an = None if anchors is None else np.ascontiguousarray(anchors, np.float32).reshape(6)
        n = self.lib.yfo_decode_nms_ex(head.ctypes.data, gh, gw, scale, zp, None if an is None else an.ctypes.data, stride,
                                       conf_thr, iou_thr, int(plus_one), dets, cap)
        return np.frombuffer(dets, np.float32, n * 5).reshape(-1, 5).copy()

    def decode_all(self, head, scale=0.14218327403068542, zp=-15, anchors=None, stride=8.0):
        """every candidate, memory order (cell-major, anchor-minor): [gh*gw*3, 5] = x1,y1,x2,y2,conf"""
        head = np.ascontiguousarray(head, dtype=np.int8)
        gh, gw, _ = head.shape
        dets = (Det * (gh * gw * 3))()
        an = None if anchors is None else np.ascontiguousarray(anchors, np.float32).reshape(6)
        self.lib.yfo_decode_all(head.ctypes.data, gh, gw, scale, zp, None if an is None else an.ctypes.data, stride, dets)
        return np.frombuffer(dets, np.float32, gh * gw * 15).reshape(-1, 5).copy()

    def rgb565_to_input(self, frame):
        frame = np.ascontiguousarray(frame, dtype=np.uint8)
        assert frame.size == 112 * 112 * 2
        out = np.empty((56, 56, 3), np.int8)
        self.lib.yfo_rgb565_to_input(frame.ctypes.data, out.ctypes.data)
        return out


# ---- the survey's known-answer inputs --------------------------------
def vector_a():
    i = np.arange(56 * 56 * 3, dtype=np.int64)
    return (((37 * i + 11) % 256) - 128).astype(np.int8).reshape(56, 56, 3)


def vector_b():
    s = 12345; x = np.empty(56 * 56 * 3, np.int8)
    for k in range(x.size):
        s = (s * 1103515245 + 12345) & 0x7FFFFFFF
        x[k] = ((s >> 16) & 0xFF) - 128
    return x.reshape(56, 56, 3)
