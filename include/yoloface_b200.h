/*
 * yoloface_b200.h -- B200 extensions next to the X-CUBE-AI style API of network.h.
 *
 * The reference API stops at "one int8 image in, one int8 head out" (ai_network_run,
 * network.h:193-195) with a 16-bit batch field (ai_platform.h:519).  These entry points cover
 * what a GPU caller needs beyond that, without changing the reference calls:
 *   - batches above 65,535 and explicit host/device pointers           (yf_b200_run)
 *   - head decode + NMS on device: yoloface.c:105-152 / tflite_prediction.py:43-57 /
 *     yoloface_test.py:148-201                                          (yf_b200_detect, yf_b200_decode)
 *   - camera-side pre-processing on device: yoloface.c:26-93           (yf_b200_preprocess_rgb565)
 *   - per-operator tensors, the analogue of ST's observer API
 *     (ai_platform_interface.h:695-731)                                 (yf_b200_set_observer, yf_b200_get_tensor)
 *   - other input resolutions of the fully-convolutional graph          (yf_b200_set_input_size)
 * All functions return >= 0 on success and a negative value on failure, latching an ai_error
 * readable through ai_network_get_error().
 */
#ifndef YOLOFACE_B200_H
#define YOLOFACE_B200_H

#include "ai_platform.h"

AI_API_DECLARE_BEGIN

#define YF_B200_CONFIG_MAGIC 0x32424659u /* "YFB2" */

/* Optional configuration, passed to ai_network_create() as
 *   ai_buffer cfg = AI_BUFFER_OBJ_INIT(AI_BUFFER_FORMAT_U8, 1, 1, sizeof(yf_b200_config), 1, &config);
 * NULL keeps the defaults (environment: YF_B200_DEVICE, YF_B200_DEVICES, YF_B200_CHUNK, YF_B200_TFLITE). */
typedef struct yf_b200_config_ {
  uint32_t magic;          /* YF_B200_CONFIG_MAGIC */
  int32_t device;          /* CUDA ordinal; -1 = YF_B200_DEVICE or the current device */
  uint32_t chunk_images;   /* images processed per pipeline chunk; 0 = default */
  uint32_t flags;          /* YF_B200_FLAG_* */
  const char* tflite_path; /* NULL = model embedded in the library */
  uint32_t device_mask;    /* bit d set = use CUDA device d.  Two or more bits make ONE handle drive several GPUs: the
                            * images of every ai_network_run / yf_b200_run / yf_b200_detect call with host buffers are
                            * split into contiguous ranges, one per device (a worker thread + streams per GPU inside the
                            * library), results land in the caller's buffers at the ranges' offsets; no collective.
                            * 0 = the single `device` above.  Read only when the ai_buffer announces a struct this large
                            * (channels = sizeof(yf_b200_config)); environment: YF_B200_DEVICES=all | 0,1,2 */
} yf_b200_config;

#define YF_B200_FLAG_OBSERVER 0x1u   /* keep every operator's tensor (slower, more memory) */
#define YF_B200_FLAG_LAYERED 0x2u    /* always run the layer-by-layer kernels (one launch per fused step) */
#define YF_B200_FLAG_ST_ACTIVATIONS 0x8u /* LeakyReLU tables as ST's code generator rounds them (network.c:2218..2902)
                                          * instead of TFLite's fixed-point rule; also env YF_B200_ST_ACTIVATIONS=1 */
#define YF_B200_FLAG_FUSED_ONLY 0x4u /* fail instead of falling back when the single-kernel path cannot be used */

typedef struct yf_b200_det_ {
  float x1, y1, x2, y2, conf;        /* corners in input pixels (tflite_prediction.py:5-11), confidence */
} yf_b200_det;

#define YF_B200_NMS_PLUS_ONE 0x1     /* integer "+1" box area convention of yoloface_test.py:172-186 */

/* Input size of the fully convolutional graph (multiples of 8); default 56x56 (network.h:38-52). */
AI_API_ENTRY int32_t yf_b200_set_input_size(ai_handle network, int32_t height, int32_t width);

/* n images [n,H,W,3] int8 -> heads [n,H/8,W/8,18] int8.  `in`/`out` may each be host or device
 * memory (host memory is fastest when page-locked).  Returns n. */
AI_API_ENTRY int32_t yf_b200_run(ai_handle network, const void* in, void* out, uint32_t n);

/* Stream integration: run on the caller's CUDA stream (cudaStream_t; NULL = the library's own).
 * yf_b200_enqueue queues the kernels for n device-resident images without synchronising (device
 * pointers only, input 16-byte aligned); yf_b200_sync waits and reports pipeline errors. */
AI_API_ENTRY int32_t yf_b200_set_stream(ai_handle network, void* cuda_stream);
AI_API_ENTRY int32_t yf_b200_enqueue(ai_handle network, const void* d_in, void* d_out, uint32_t n);
AI_API_ENTRY int32_t yf_b200_sync(ai_handle network);
/* Several independent device-resident batches in one call (batch b: counts[b] images at d_in[b] ->
 * d_out[b]).  The batches are ordered after earlier work on the stream and before later work, but not
 * among themselves: the library spreads them over four internal kernel lanes, so the first CTAs of one
 * batch fill the SM slots the previous batches leave idle (one 256-image launch occupies 256 of the 444
 * resident-CTA slots).  yf_b200_enqueue does the same for the chunks of one large batch.  Returns the
 * number of images queued. */
AI_API_ENTRY int32_t yf_b200_enqueue_batches(ai_handle network, const void* const* d_in, void* const* d_out,
                                             const uint32_t* counts, uint32_t n_batches);

/* Pipelined host path: yf_b200_submit queues n images from (preferably page-locked) host memory --
 * H2D copy, kernels (alternating over four lanes) and D2H copy of the heads run on separate streams over a
 * ring of staging slots, so the copies of one submission overlap the kernels of its neighbours -- and
 * returns without waiting;
 * yf_b200_wait blocks until every submitted result is in `out_host` and reports pipeline errors.
 * (yf_b200_run / ai_network_run use the same ring internally for batches larger than one chunk.) */
AI_API_ENTRY int32_t yf_b200_submit(ai_handle network, const void* in_host, void* out_host, uint32_t n);
AI_API_ENTRY int32_t yf_b200_wait(ai_handle network);

/* Decode + NMS of heads already computed ([n,gh,gw,18] int8, host or device).
 * conf_thr: keep conf >= conf_thr (0.7 in yoloface.c:123); iou_thr < 0: threshold only (what the
 * firmware does), else greedy NMS keeping iou <= iou_thr (0.4 in yoloface_test.py:32).
 * dets: [n, max_det] records, counts: [n] (host memory).  Returns total detections. */
AI_API_ENTRY int32_t yf_b200_decode(ai_handle network, const void* heads, uint32_t n, float conf_thr, float iou_thr,
                                    uint32_t flags, yf_b200_det* dets, int32_t* counts, uint32_t max_det);

/* Anchor table (3 x {w, h} in input pixels; NULL keeps the current one, default = yoloface.c:20) and the input
 * pixels per head cell (0 = input height / head rows, 8 for this model: yoloface.c:135-136) used by the decode.
 * Every candidate that passes conf_thr takes part in the NMS, whatever the head size: heads up to 8x8 cells use one
 * warp per image, larger ones one thread block per image with storage for all gh*gw*3 candidates (heads beyond
 * 52x52 cells are refused with an error, never truncated). */
AI_API_ENTRY int32_t yf_b200_set_decode_params(ai_handle network, const float* anchors6, float stride);

/* Inference + decode + NMS in one call; only detections travel back to the host.
 * heads_out may be NULL. */
AI_API_ENTRY int32_t yf_b200_detect(ai_handle network, const void* in, uint32_t n, float conf_thr, float iou_thr,
                                    uint32_t flags, yf_b200_det* dets, int32_t* counts, uint32_t max_det,
                                    void* heads_out);

/* n RGB565 112x112 frames (big-endian byte pairs, as OV2640 delivers them: yoloface.c:41-47)
 * -> int8 [n,56,56,3] network inputs.  Host or device pointers. */
AI_API_ENTRY int32_t yf_b200_preprocess_rgb565(ai_handle network, const void* frames, void* out, uint32_t n);

/* Observer mode: after a run of at most one chunk, every TFLite tensor that is materialised can
 * be read back densely ([n,H,W,C] int8).  Returns bytes written, or <0 if that tensor is folded
 * away (PAD outputs) or n exceeds the last run. */
AI_API_ENTRY int32_t yf_b200_set_observer(ai_handle network, int32_t enable);
AI_API_ENTRY int64_t yf_b200_get_tensor(ai_handle network, int32_t tflite_tensor, uint32_t n, void* dst, uint64_t dst_bytes);
/* shape of a TFLite tensor at the current input size: dims[4] = {1,H,W,C}; returns 0 if known */
AI_API_ENTRY int32_t yf_b200_tensor_shape(ai_handle network, int32_t tflite_tensor, int32_t dims[4]);

typedef struct yf_b200_stats_ {
  uint64_t kernel_launches;   /* kernels of this library launched since create */
  uint64_t images;            /* images inferred since create */
  float last_run_device_ms;   /* CUDA-event time of the last run's kernels (excl. copies); 0 after a pipelined host-to-host call */
  int32_t device;
  int32_t sm_count;
  uint32_t chunk_images;
  int32_t steps;              /* fused device steps per image batch (26 for yoloface) */
  int32_t fused;              /* 1: runs go through the single persistent kernel; 0: layer-by-layer */
  int32_t fused_smem_bytes;   /* dynamic shared memory of the fused kernel */
  int32_t fused_latency;      /* 1: launches of at most one image per SM that run alone use the latency shape (512-thread CTAs) */
  uint32_t latency_launches;  /* how many launches took that shape since create */
  int32_t cluster_images;     /* > 0: launches of at most this many images that run alone give every image a CLUSTER of 4 such CTAs */
  uint32_t cluster_launches;  /* how many launches did */
} yf_b200_stats;
AI_API_ENTRY int32_t yf_b200_get_stats(ai_handle network, yf_b200_stats* stats);

/* Per-step description (name, TFLite ops folded, algorithmic bytes/MACs per image) for reports. */
typedef struct yf_b200_step_info_ {
  char name[32];
  int32_t kind;               /* 0 im2col conv, 1 conv1x1, 2 depthwise, 3 maxpool, 4 lut */
  int32_t first_op, n_ops;
  int64_t macs;               /* per image */
  int64_t bytes_read, bytes_written;  /* algorithmic (unpadded) per image, weights excluded */
  float last_ms;              /* device time of this step in the last profiled run, <0 if none */
} yf_b200_step_info;
AI_API_ENTRY int32_t yf_b200_step_count(ai_handle network);
AI_API_ENTRY int32_t yf_b200_step_info_get(ai_handle network, int32_t step, yf_b200_step_info* info);
/* time every step of the next run with CUDA events (adds a sync per step) */
AI_API_ENTRY int32_t yf_b200_set_step_profiling(ai_handle network, int32_t enable);

/* Cycle trace of the fused kernel: enable, run once, then read nphases+1 SM-clock stamps taken by
 * CTA 0 at every phase boundary of its first image (profiling aid). */
AI_API_ENTRY int32_t yf_b200_fused_trace(ai_handle network, int32_t enable, int64_t* stamps, int32_t cap);

/* Page-locked host memory for the fast host path. */
AI_API_ENTRY void* yf_b200_host_alloc(uint64_t bytes);
AI_API_ENTRY void yf_b200_host_free(void* p);

/* Plan introspection -- pure host code, usable without a GPU (tests validate the lowering on CPU).
 * yf_b200_plan_json: JSON description of the fused steps / buffers / tensor placement for an HxW
 * input built from the embedded model (weights from `blob` in ST layout, or the model's own if
 * NULL).  yf_b200_plan_blob: raw tables, what = 0 EpiCh[], 1 LUTs (n x 256), 2 packed weights.
 * Both return the number of bytes needed (copying min(cap, needed)), <0 on failure. */
AI_API_ENTRY int64_t yf_b200_plan_json(int32_t height, int32_t width, const void* blob, char* dst, uint64_t cap);
/* same for the fused single-kernel program (smem map + phases); plan_blob what = 3: its parameter
 * blob, 4: its EpiCh table.  <0 if this input size cannot run fused. */
AI_API_ENTRY int64_t yf_b200_fused_json(int32_t height, int32_t width, const void* blob, char* dst, uint64_t cap);
/* the same program laid out for a CTA of `threads` threads: 256 = throughput shape (what yf_b200_fused_json describes),
 * 512 = latency shape (launches of at most one image per SM; plan_blob what = 5 is its parameter blob),
 * 4512 = cluster shape (512-thread CTAs in clusters of 4 per image; plan_blob what = 6). */
AI_API_ENTRY int64_t yf_b200_fused_json_ex(int32_t height, int32_t width, const void* blob, int32_t threads, char* dst, uint64_t cap);
AI_API_ENTRY int64_t yf_b200_plan_blob(int32_t height, int32_t width, const void* blob, int32_t what, void* dst, uint64_t cap);

/* Human-readable text of the last failure (CUDA error string, plan error, ...). */
AI_API_ENTRY const char* yf_b200_last_error_text(void);

/* Test hook: queue a one-thread kernel that sets the pipeline error word the way a kernel whose bounded wait gave up
 * does; the next synchronising call (yf_b200_sync / yf_b200_wait / a run) must fail with AI_ERROR_INVALID_STATE /
 * AI_ERROR_CODE_LAYER and clear the word. */
AI_API_ENTRY int32_t yf_b200_debug_raise(ai_handle network, int32_t code);

AI_API_DECLARE_END
#endif
