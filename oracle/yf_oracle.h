/*
 * yf_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the TensorFlow-Lite reference int8 kernels needed to
 * run `yoloface/tflite/yoloface_int8.tflite`, the model the reference runs
 * through `tf.lite.Interpreter` (yoloface/tflite/tflite_prediction.py:23-41)
 * and that X-CUBE-AI compiles into stm32/X-CUBE-AI/App/network.c.
 *
 * PARITY STATUS: "parity unpinned" against a live TFLite interpreter -- the
 * reference repository holds no recorded outputs for this path and TensorFlow
 * (pinned tensorflow==2.10.0, yoloface/tensorflow/requirements.txt:2) is not
 * installable here.  What pins the oracle instead (tests/test_oracle_*.py):
 *   - the survey's independent numpy restatement (per-op CRC32s),
 *   - the 17 ST LeakyReLU LUTs in network.c (must differ in exactly 271 entries, by 1 LSB),
 *   - byte-identity of the weight blob in network_data.c at the offsets of network.c:3117-3263,
 *   - the quantisation constants duplicated in network.c:663-1341 / yoloface.c:116.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (libyoloface_b200.so) never links it.
 */
#ifndef YF_ORACLE_H
#define YF_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* TFLite builtin operator codes used by the model (schema.fbs BuiltinOperator). */
enum {
  YFO_OP_ADD = 0, YFO_OP_CONCATENATION = 2, YFO_OP_CONV_2D = 3, YFO_OP_DEPTHWISE_CONV_2D = 4,
  YFO_OP_DEQUANTIZE = 6, YFO_OP_MAX_POOL_2D = 17, YFO_OP_PAD = 34, YFO_OP_LEAKY_RELU = 98,
  YFO_OP_QUANTIZE = 114
};

typedef struct yfo_model yfo_model;

/* Parse a .tflite flatbuffer (the bytes must stay alive for the model's life). NULL on error. */
yfo_model* yfo_load(const uint8_t* buf, size_t len);
void yfo_free(yfo_model* m);
const char* yfo_last_error(void);

int yfo_num_tensors(const yfo_model* m);
int yfo_num_ops(const yfo_model* m);
int yfo_input_tensor(const yfo_model* m);
int yfo_output_tensor(const yfo_model* m);
/* shape: up to 4 dims (returns rank); type: TFLite TensorType (9=int8, 2=int32); nscale = #scales */
int yfo_tensor_info(const yfo_model* m, int t, int shape[4], int* type, int* nscale, int* qdim,
                    const uint8_t** data, size_t* data_len);
float yfo_tensor_scale(const yfo_model* m, int t, int i);
int64_t yfo_tensor_zp(const yfo_model* m, int t, int i);
const char* yfo_tensor_name(const yfo_model* m, int t);
/* opcode + up to 3 inputs / 1 output tensor indices; returns #inputs */
int yfo_op_info(const yfo_model* m, int op, int* opcode, int inputs[3], int* output);

/* Output element count of op `op` when the network input is [1,H,W,3] (shapes are propagated,
 * the graph is fully convolutional: yoloface/tensorflow/yolo_to_h5.py:134). <0 on error. */
long yfo_op_out_elems(const yfo_model* m, int op, int H, int W, int shape_out[4]);

/* Run one image.  in: int8 [H,W,3].  out: int8 [H/8,W/8,18].  op_out: NULL, or an array of
 * yfo_num_ops() pointers (entries may be NULL) receiving every op's output tensor (NHWC int8). */
int yfo_run(const yfo_model* m, const int8_t* in, int H, int W, int8_t* out, int8_t** op_out);
/* Run n images with `threads` worker threads (pthreads); returns 0 on success. */
int yfo_run_batch(const yfo_model* m, const int8_t* in, int n, int H, int W, int8_t* out, int threads);

/* ---- fixed-point primitives (TFLite common.h / gemmlowp fixedpoint.h, double-rounding build) ---- */
void yfo_quantize_multiplier(double d, int32_t* mult, int* shift);
int32_t yfo_srdhm(int32_t a, int32_t b);
int32_t yfo_rdivpot(int32_t x, int exponent);
int32_t yfo_mbqm(int32_t x, int32_t mult, int shift);

/* 256-entry int8->int8 table of LEAKY_RELU op `op` with TFLite arithmetic: lut[q+128]. */
int yfo_leaky_lut(const yfo_model* m, int op, int8_t lut[256]);

/* ---- head decode + NMS (yoloface.c:105-152, tflite_prediction.py:43-57, yoloface_test.py:148-201) ---- */
typedef struct { float x1, y1, x2, y2, conf; } yfo_det;
/* head: int8 [gh,gw,18]; candidates in memory order (cell-major, anchor-minor, as yoloface.c:109-116).
 * Keeps conf >= conf_thr, greedy NMS keeping iou <= iou_thr (iou_thr<0: threshold only, no NMS).
 * plus_one!=0 selects the integer "+1" area convention of yoloface_test.py:172-186.
 * Returns #detections written (<= max_det), sorted by conf desc, ties by candidate index asc. */
int yfo_decode_nms(const int8_t* head, int gh, int gw, float out_scale, int out_zp,
                   float conf_thr, float iou_thr, int plus_one, yfo_det* dets, int max_det);
/* The same with the anchor table (3 x {w,h}, NULL = yoloface.c:20) and the cell stride as parameters. */
int yfo_decode_nms_ex(const int8_t* head, int gh, int gw, float out_scale, int out_zp, const float* anchors6, float stride,
                      float conf_thr, float iou_thr, int plus_one, yfo_det* dets, int max_det);
/* Every candidate decoded, no threshold, memory order (cands[gh*gw*3], index = cell*3 + anchor). */
void yfo_decode_all(const int8_t* head, int gh, int gw, float out_scale, int out_zp, const float* anchors6, float stride,
                    yfo_det* cands);

/* ---- camera-side pre-processing (yoloface.c:26-93): RGB565 112x112 big-endian byte pairs ->
 * 2x2 box average per 5/6/5 field -> expand (r<<3,g<<2,b<<3) -> -128 -> int8 [56,56,3] ---- */
void yfo_rgb565_to_input(const uint8_t* rgb565_112, int8_t* in_56x56x3);

#ifdef __cplusplus
}
#endif
#endif
