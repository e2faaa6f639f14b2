"""stm32h7_yolo_b200 -- host-side mirror of the reference's caller code for the yoloface int8 path.

The product is libyoloface_b200.so (C ABI: include/network.h, include/network_data.h,
include/yoloface_b200.h).  This module is the thin ctypes binding a Python caller uses in place of
`tf.lite.Interpreter` (yoloface/tflite/tflite_prediction.py:23-41 of the reference); the C caller
equivalent of yoloface.c's aiInit/aiRun is host/yoloface_app.c.

There is no CPU fallback here: if the shared library is missing or no B200 is present, loading
or `Network()` raises.
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("YF_B200_LIB") or os.path.join(PKG_DIR, "libyoloface_b200.so")   # YF_B200_LIB: e.g. the TRACE=1 twin

# ---- ABI structs (include/ai_platform.h) ------------------------------------------------------
AI_BUFFER_FORMAT_S8 = (1 << 23) | (2 << 17) | (8 << 7) | 64
AI_BUFFER_FORMAT_U8 = (2 << 17) | (8 << 7) | 64
AI_BUFFER_FMT_FLAG_CONST = 1 << 30
AI_NETWORK_DATA_WEIGHTS_SIZE = 11304
AI_NETWORK_DATA_ACTIVATIONS_SIZE = 29784
YF_B200_CONFIG_MAGIC = 0x32424659
YF_B200_FLAG_OBSERVER = 1
YF_B200_FLAG_LAYERED = 2
YF_B200_FLAG_FUSED_ONLY = 4
YF_B200_FLAG_ST_ACTIVATIONS = 8
YF_B200_NMS_PLUS_ONE = 1


class AiBuffer(C.Structure):
    _fields_ = [("format", C.c_int32), ("n_batches", C.c_uint16), ("height", C.c_uint16), ("width", C.c_uint16),
                ("channels", C.c_uint32), ("data", C.c_void_p), ("meta_info", C.c_void_p)]


class AiError(C.Structure):
    _fields_ = [("type", C.c_uint32, 8), ("code", C.c_uint32, 24)]


class AiNetworkParams(C.Structure):
    _fields_ = [("params", AiBuffer), ("activations", AiBuffer)]


class Config(C.Structure):
    _fields_ = [("magic", C.c_uint32), ("device", C.c_int32), ("chunk_images", C.c_uint32), ("flags", C.c_uint32),
                ("tflite_path", C.c_char_p), ("device_mask", C.c_uint32)]


class Det(C.Structure):
    _fields_ = [("x1", C.c_float), ("y1", C.c_float), ("x2", C.c_float), ("y2", C.c_float), ("conf", C.c_float)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("images", C.c_uint64), ("last_run_device_ms", C.c_float),
                ("device", C.c_int32), ("sm_count", C.c_int32), ("chunk_images", C.c_uint32), ("steps", C.c_int32),
                ("fused", C.c_int32), ("fused_smem_bytes", C.c_int32), ("fused_latency", C.c_int32), ("latency_launches", C.c_uint32),
                ("cluster_images", C.c_int32), ("cluster_launches", C.c_uint32)]


class StepInfo(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("kind", C.c_int32), ("first_op", C.c_int32), ("n_ops", C.c_int32),
                ("macs", C.c_int64), ("bytes_read", C.c_int64), ("bytes_written", C.c_int64), ("last_ms", C.c_float)]


assert C.sizeof(AiBuffer) == 32 and C.sizeof(AiError) == 4 and C.sizeof(AiNetworkParams) == 64

EXPORTS = [
    "ai_network_create", "ai_network_init", "ai_network_run", "ai_network_forward", "ai_network_destroy",
    "ai_network_get_error", "ai_network_get_report", "ai_network_get_info", "ai_network_data_weights_get",
    "ai_network_data_params_get", "yf_b200_set_input_size", "yf_b200_run", "yf_b200_decode", "yf_b200_detect",
    "yf_b200_preprocess_rgb565", "yf_b200_set_decode_params", "yf_b200_set_observer", "yf_b200_get_tensor", "yf_b200_tensor_shape",
    "yf_b200_get_stats", "yf_b200_step_count", "yf_b200_step_info_get", "yf_b200_set_step_profiling",
    "yf_b200_fused_trace", "yf_b200_submit", "yf_b200_wait", "yf_b200_set_stream", "yf_b200_enqueue", "yf_b200_enqueue_batches", "yf_b200_sync", "yf_b200_host_alloc", "yf_b200_host_free", "yf_b200_last_error_text", "yf_b200_debug_raise", "yf_b200_plan_json", "yf_b200_plan_blob", "yf_b200_fused_json", "yf_b200_fused_json_ex",
]

_lib = None


def build(force=False):
    """Compile libyoloface_b200.so in-tree (nvcc, sm_100a)."""
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-s", "-C", os.path.join(PKG_DIR, "csrc")])
    return LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libyoloface_b200.so is not built (run __graft_entry__.build()); there is no fallback path")
    L = C.CDLL(LIB_PATH)
    vp, i32, u32 = C.c_void_p, C.c_int32, C.c_uint32
    L.ai_network_create.restype = AiError
    L.ai_network_create.argtypes = [C.POINTER(vp), C.POINTER(AiBuffer)]
    L.ai_network_init.restype = C.c_bool
    L.ai_network_init.argtypes = [vp, C.POINTER(AiNetworkParams)]
    L.ai_network_run.restype = i32
    L.ai_network_run.argtypes = [vp, C.POINTER(AiBuffer), C.POINTER(AiBuffer)]
    L.ai_network_forward.restype = i32
    L.ai_network_forward.argtypes = [vp, C.POINTER(AiBuffer)]
    L.ai_network_destroy.restype = vp
    L.ai_network_destroy.argtypes = [vp]
    L.ai_network_get_error.restype = AiError
    L.ai_network_get_error.argtypes = [vp]
    L.ai_network_data_weights_get.restype = vp
    L.ai_network_data_weights_get.argtypes = []
    L.ai_network_data_params_get.restype = C.c_bool
    L.ai_network_data_params_get.argtypes = [vp, vp]
    L.yf_b200_set_input_size.restype = i32
    L.yf_b200_set_input_size.argtypes = [vp, i32, i32]
    L.yf_b200_run.restype = i32
    L.yf_b200_run.argtypes = [vp, vp, vp, u32]
    L.yf_b200_fused_trace.restype = i32
    L.yf_b200_fused_trace.argtypes = [vp, i32, C.POINTER(C.c_int64), i32]
    L.yf_b200_set_stream.restype = i32
    L.yf_b200_set_stream.argtypes = [vp, vp]
    L.yf_b200_enqueue.restype = i32
    L.yf_b200_enqueue.argtypes = [vp, vp, vp, u32]
    L.yf_b200_enqueue_batches.restype = i32
    L.yf_b200_enqueue_batches.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(u32), u32]
    L.yf_b200_submit.restype = i32
    L.yf_b200_submit.argtypes = [vp, vp, vp, u32]
    L.yf_b200_wait.restype = i32
    L.yf_b200_wait.argtypes = [vp]
    L.yf_b200_sync.restype = i32
    L.yf_b200_sync.argtypes = [vp]
    L.yf_b200_decode.restype = i32
    L.yf_b200_decode.argtypes = [vp, vp, u32, C.c_float, C.c_float, u32, vp, vp, u32]
    L.yf_b200_detect.restype = i32
    L.yf_b200_detect.argtypes = [vp, vp, u32, C.c_float, C.c_float, u32, vp, vp, u32, vp]
    L.yf_b200_preprocess_rgb565.restype = i32
    L.yf_b200_preprocess_rgb565.argtypes = [vp, vp, vp, u32]
    L.yf_b200_set_decode_params.restype = i32
    L.yf_b200_set_decode_params.argtypes = [vp, vp, C.c_float]
    L.yf_b200_set_observer.restype = i32
    L.yf_b200_set_observer.argtypes = [vp, i32]
    L.yf_b200_get_tensor.restype = C.c_int64
    L.yf_b200_get_tensor.argtypes = [vp, i32, u32, vp, C.c_uint64]
    L.yf_b200_tensor_shape.restype = i32
    L.yf_b200_tensor_shape.argtypes = [vp, i32, C.POINTER(i32)]
    L.yf_b200_get_stats.restype = i32
    L.yf_b200_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.yf_b200_step_count.restype = i32
    L.yf_b200_step_count.argtypes = [vp]
    L.yf_b200_step_info_get.restype = i32
    L.yf_b200_step_info_get.argtypes = [vp, i32, C.POINTER(StepInfo)]
    L.yf_b200_set_step_profiling.restype = i32
    L.yf_b200_set_step_profiling.argtypes = [vp, i32]
    L.yf_b200_host_alloc.restype = vp
    L.yf_b200_host_alloc.argtypes = [C.c_uint64]
    L.yf_b200_host_free.argtypes = [vp]
    L.yf_b200_last_error_text.restype = C.c_char_p
    L.yf_b200_plan_json.restype = C.c_int64
    L.yf_b200_plan_json.argtypes = [i32, i32, vp, C.c_char_p, C.c_uint64]
    L.yf_b200_fused_json.restype = C.c_int64
    L.yf_b200_fused_json.argtypes = [i32, i32, vp, C.c_char_p, C.c_uint64]
    L.yf_b200_fused_json_ex.restype = C.c_int64
    L.yf_b200_fused_json_ex.argtypes = [i32, i32, vp, i32, C.c_char_p, C.c_uint64]
    L.yf_b200_plan_blob.restype = C.c_int64
    L.yf_b200_plan_blob.argtypes = [i32, i32, vp, i32, vp, C.c_uint64]
    _lib = L
    return L


class AiRuntimeError(RuntimeError):
    def __init__(self, what, err, text):
        super().__init__("%s failed: ai_error{type=0x%02x, code=0x%04x} %s" % (what, err.type, err.code, text))
        self.type, self.code = err.type, err.code


def _ptr(x):
    """numpy array / torch tensor / int address -> (address, keepalive)."""
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"]
        return x.ctypes.data, x
    if hasattr(x, "data_ptr"):
        assert x.is_contiguous()
        return x.data_ptr(), x
    return int(x), None


def weights_blob():
    """The 11,304-byte ST-layout blob behind ai_network_data_weights_get() (no GPU needed)."""
    L = lib()
    table = C.cast(L.ai_network_data_weights_get(), C.POINTER(C.c_void_p))
    assert table[0] == 0xA1FACADE and table[2] == 0xA1FACADE
    return C.string_at(table[1], AI_NETWORK_DATA_WEIGHTS_SIZE)


def plan(height=56, width=56, blob=None):
    """Host-side plan description + tables (no GPU needed)."""
    L = lib()
    bp = C.create_string_buffer(blob, len(blob)) if blob is not None else None
    n = L.yf_b200_plan_json(height, width, bp, None, 0)
    if n < 0:
        raise RuntimeError(L.yf_b200_last_error_text().decode())
    buf = C.create_string_buffer(n)
    L.yf_b200_plan_json(height, width, bp, buf, n)
    desc = json.loads(buf.value.decode())
    tables = []
    for what in range(3):
        k = L.yf_b200_plan_blob(height, width, bp, what, None, 0)
        raw = C.create_string_buffer(max(int(k), 1))
        L.yf_b200_plan_blob(height, width, bp, what, raw, k)
        tables.append(raw.raw[:k])
    desc["epi"] = np.frombuffer(tables[0], dtype=np.dtype([("add64", "<i8"), ("mult", "<i4"), ("c2", "<i4"), ("e", "<i4"),
                                                            ("ls", "<i4"), ("sgn_mask", "<i4"), ("acc_bound", "<i4")]))
    desc["luts"] = np.frombuffer(tables[1], dtype=np.int8).reshape(-1, 256)
    desc["wblob"] = np.frombuffer(tables[2], dtype=np.uint8)
    return desc


EPI_DTYPE = np.dtype([("add64", "<i8"), ("mult", "<i4"), ("c2", "<i4"), ("e", "<i4"), ("ls", "<i4"), ("sgn_mask", "<i4"), ("acc_bound", "<i4")])


def fused_program(height=56, width=56, blob=None, threads=256):
    """Fused single-kernel program (smem map, phases, parameter blob, EpiCh table); no GPU needed.
    threads: CTA shape the program is laid out for (256 throughput, 512 latency, 4512 latency with clusters of 4)."""
    L = lib()
    bp = C.create_string_buffer(blob, len(blob)) if blob is not None else None
    n = L.yf_b200_fused_json_ex(height, width, bp, threads, None, 0)
    if n < 0:
        raise RuntimeError(L.yf_b200_last_error_text().decode())
    buf = C.create_string_buffer(n)
    L.yf_b200_fused_json_ex(height, width, bp, threads, buf, n)
    prog = json.loads(buf.value.decode())
    for key, what in (("params", {512: 5, 4512: 6}.get(threads, 3)), ("epi", 4)):
        k = L.yf_b200_plan_blob(height, width, bp, what, None, 0)
        raw = C.create_string_buffer(max(int(k), 1))
        L.yf_b200_plan_blob(height, width, bp, what, raw, k)
        prog[key] = raw.raw[:k]
    prog["params"] = np.frombuffer(prog["params"], dtype=np.uint8)
    prog["epi"] = np.frombuffer(prog["epi"], dtype=EPI_DTYPE)
    return prog


class Network:
    """aiInit()/aiRun() of stm32/X-CUBE-AI/App/yoloface.c:188-240, batch-capable."""

    def __init__(self, device=-1, chunk_images=0, observer=False, tflite_path=None, weights=None, mode="auto", st_activations=False,
                 devices=None):
        """devices: list of CUDA ordinals driven by this ONE handle (batches of run()/detect()/ai_run() with host
        buffers are split by image over them); None = the single `device`."""
        L = self.L = lib()
        self.handle = C.c_void_p()
        flags = (YF_B200_FLAG_OBSERVER if observer else 0) | {"auto": 0, "layered": YF_B200_FLAG_LAYERED,
                                                              "fused": YF_B200_FLAG_FUSED_ONLY}[mode]
        flags |= YF_B200_FLAG_ST_ACTIVATIONS if st_activations else 0
        self._cfg = Config(YF_B200_CONFIG_MAGIC, device, chunk_images, flags,
                           tflite_path.encode() if tflite_path else None, sum(1 << d for d in devices) if devices else 0)
        cfgbuf = AiBuffer(AI_BUFFER_FORMAT_U8, 1, 1, 1, C.sizeof(Config), C.cast(C.pointer(self._cfg), C.c_void_p), None)
        err = L.ai_network_create(C.byref(self.handle), C.byref(cfgbuf))
        if err.type != 0:
            raise AiRuntimeError("ai_network_create", err, L.yf_b200_last_error_text().decode())
        # yoloface.c:198-201: weights handle from network_data, caller-owned activations arena
        self._activations = (C.c_uint8 * AI_NETWORK_DATA_ACTIVATIONS_SIZE)()
        if weights is None and tflite_path:
            wptr = None                                # a model given by path keeps the weights of its own flatbuffer
        elif weights is None:
            wptr = L.ai_network_data_weights_get()
        else:
            self._weights = C.create_string_buffer(bytes(weights), len(weights))
            wptr = C.cast(self._weights, C.c_void_p).value
        params = AiNetworkParams(
            AiBuffer(AI_BUFFER_FORMAT_U8 | AI_BUFFER_FMT_FLAG_CONST, 1, 1, 1, AI_NETWORK_DATA_WEIGHTS_SIZE, wptr, None),
            AiBuffer(AI_BUFFER_FORMAT_U8, 1, 1, 1, AI_NETWORK_DATA_ACTIVATIONS_SIZE,
                     C.cast(self._activations, C.c_void_p), None))
        if not L.ai_network_init(self.handle, C.byref(params)):
            self._raise("ai_network_init")
        self.H, self.W = 56, 56

    def _raise(self, what):
        err = self.L.ai_network_get_error(self.handle)
        raise AiRuntimeError(what, err, self.L.yf_b200_last_error_text().decode())

    def get_error(self):
        e = self.L.ai_network_get_error(self.handle)
        return e.type, e.code

    def close(self):
        if self.handle:
            self.L.ai_network_destroy(self.handle)
            self.handle = C.c_void_p()

    def set_input_size(self, H, W):
        if self.L.yf_b200_set_input_size(self.handle, H, W) < 0:
            self._raise("yf_b200_set_input_size")
        self.H, self.W = H, W

    # ---- the reference call: ai_network_run with ai_buffer descriptors (yoloface.c:216-240) ----
    def ai_run(self, inp, out=None):
        n = inp.shape[0]
        if out is None:
            out = np.empty((n, self.H // 8, self.W // 8, 18), np.int8)
        ip, _k1 = _ptr(inp)
        op, _k2 = _ptr(out)
        bi = AiBuffer(AI_BUFFER_FORMAT_S8, n, self.H, self.W, 3, ip, None)
        bo = AiBuffer(AI_BUFFER_FORMAT_S8, n, self.H // 8, self.W // 8, 18, op, None)
        r = self.L.ai_network_run(self.handle, C.byref(bi), C.byref(bo))
        if r != n:
            self._raise("ai_network_run")
        return out

    # ---- extensions -------------------------------------------------------------------------
    def run(self, inp, out=None, n=None):
        """inp: numpy int8 [n,H,W,3] (host) or torch CUDA/pinned tensor; returns heads [n,H/8,W/8,18]."""
        if n is None:
            n = inp.shape[0]
        if out is None:
            out = np.empty((n, self.H // 8, self.W // 8, 18), np.int8)
        ip, _k1 = _ptr(inp)
        op, _k2 = _ptr(out)
        if self.L.yf_b200_run(self.handle, ip, op, n) != n:
            self._raise("yf_b200_run")
        return out

    def fused_trace(self, enable=True, read=False):
        buf = (C.c_int64 * 128)()
        k = self.L.yf_b200_fused_trace(self.handle, int(enable), buf if read else None, 128 if read else 0)
        return list(buf)[:k] if read else None

    def set_stream(self, cuda_stream):
        """cuda_stream: integer cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream) or None."""
        if self.L.yf_b200_set_stream(self.handle, cuda_stream) < 0:
            self._raise("yf_b200_set_stream")

    def enqueue(self, d_in, d_out, n):
        """Queue n device-resident images without synchronising (torch CUDA tensors or addresses)."""
        if self.L.yf_b200_enqueue(self.handle, _ptr(d_in)[0], _ptr(d_out)[0], n) != n:
            self._raise("yf_b200_enqueue")

    def enqueue_batches(self, d_ins, d_outs, counts):
        """Queue several independent device-resident batches; they may overlap each other on the GPU."""
        k = len(counts)
        ins = (C.c_void_p * k)(*[_ptr(x)[0] for x in d_ins])
        outs = (C.c_void_p * k)(*[_ptr(x)[0] for x in d_outs])
        cnt = (C.c_uint32 * k)(*[int(c) for c in counts])
        if self.L.yf_b200_enqueue_batches(self.handle, ins, outs, cnt, k) != sum(int(c) for c in counts):
            self._raise("yf_b200_enqueue_batches")

    def submit(self, in_host, out_host, n):
        """Queue n images from host memory (H2D, kernels, D2H pipelined over a ring); pair with wait()."""
        if self.L.yf_b200_submit(self.handle, _ptr(in_host)[0], _ptr(out_host)[0], n) != n:
            self._raise("yf_b200_submit")

    def wait(self):
        if self.L.yf_b200_wait(self.handle) < 0:
            self._raise("yf_b200_wait")

    def sync(self):
        if self.L.yf_b200_sync(self.handle) < 0:
            self._raise("yf_b200_sync")

    def detect(self, inp, conf_thr=0.7, iou_thr=0.4, plus_one=False, max_det=32, heads_out=None, n=None):
        if n is None:
            n = inp.shape[0]
        dets = np.zeros((n, max_det, 5), np.float32)
        counts = np.zeros(n, np.int32)
        ip, _k = _ptr(inp)
        hp = _ptr(heads_out)[0] if heads_out is not None else None
        r = self.L.yf_b200_detect(self.handle, ip, n, conf_thr, iou_thr, YF_B200_NMS_PLUS_ONE if plus_one else 0,
                                  dets.ctypes.data, counts.ctypes.data, max_det, hp)
        if r < 0:
            self._raise("yf_b200_detect")
        return dets, counts

    def decode(self, heads, conf_thr=0.7, iou_thr=0.4, plus_one=False, max_det=32):
        n = heads.shape[0]
        dets = np.zeros((n, max_det, 5), np.float32)
        counts = np.zeros(n, np.int32)
        hp, _k = _ptr(heads)
        r = self.L.yf_b200_decode(self.handle, hp, n, conf_thr, iou_thr, YF_B200_NMS_PLUS_ONE if plus_one else 0,
                                  dets.ctypes.data, counts.ctypes.data, max_det)
        if r < 0:
            self._raise("yf_b200_decode")
        return dets, counts

    def set_decode_params(self, anchors=None, stride=0.0):
        """anchors: 3 x (w, h) in input pixels (None keeps the current table); stride 0 = input height / head rows"""
        an = None if anchors is None else np.ascontiguousarray(anchors, np.float32).reshape(6)
        if self.L.yf_b200_set_decode_params(self.handle, None if an is None else an.ctypes.data, float(stride)) < 0:
            self._raise("yf_b200_set_decode_params")

    def preprocess_rgb565(self, frames):
        frames = np.ascontiguousarray(frames, np.uint8).reshape(-1, 112 * 112 * 2)
        out = np.empty((frames.shape[0], 56, 56, 3), np.int8)
        if self.L.yf_b200_preprocess_rgb565(self.handle, frames.ctypes.data, out.ctypes.data, frames.shape[0]) < 0:
            self._raise("yf_b200_preprocess_rgb565")
        return out

    def set_observer(self, on=True):
        if self.L.yf_b200_set_observer(self.handle, int(on)) < 0:
            self._raise("yf_b200_set_observer")

    def get_tensor(self, t, n):
        dims = (C.c_int32 * 4)()
        if self.L.yf_b200_tensor_shape(self.handle, t, dims) < 0:
            return None
        out = np.empty((n, dims[1], dims[2], dims[3]), np.int8)
        if self.L.yf_b200_get_tensor(self.handle, t, n, out.ctypes.data, out.nbytes) < 0:
            self._raise("yf_b200_get_tensor")
        return out

    def stats(self):
        s = Stats()
        self.L.yf_b200_get_stats(self.handle, C.byref(s))
        return {f[0]: getattr(s, f[0]) for f in Stats._fields_}

    def steps(self):
        out = []
        for i in range(self.L.yf_b200_step_count(self.handle)):
            si = StepInfo()
            self.L.yf_b200_step_info_get(self.handle, i, C.byref(si))
            d = {f[0]: getattr(si, f[0]) for f in StepInfo._fields_}
            d["name"] = d["name"].decode()
            out.append(d)
        return out

    def set_step_profiling(self, on=True):
        self.L.yf_b200_set_step_profiling(self.handle, int(on))
