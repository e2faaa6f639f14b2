/*
 * yoloface_app.c -- host-side C caller of the B200 library, the counterpart of the reference's
 * stm32/X-CUBE-AI/App/yoloface.c (aiInit :188-211, aiRun :216-240, post_process :105-152) and of its
 * frame loop stm32/User/main.c:42-54, but for a batch of frames and with decode + NMS on device.
 *
 *   yoloface_app [n_images] [conf_thr] [iou_thr]
 *
 * Reads nothing from disk: it synthesises RGB565 112x112 "camera frames", runs the device-side
 * pre-processing (yoloface.c:26-93), the int8 network and the decode, and prints detections in the
 * firmware's UART format (main.c:46,53 / yoloface.c:148) so the reference's PC monitor could parse them.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "network.h"
#include "network_data.h"
#include "yoloface_b200.h"

static ai_handle network = AI_HANDLE_NULL;
AI_ALIGNED(32) static ai_u8 activations[AI_NETWORK_DATA_ACTIVATIONS_SIZE];

static int aiInit(void) {
  ai_error err = ai_network_create(&network, AI_NETWORK_DATA_CONFIG);
  if (err.type != AI_ERROR_NONE) {
    printf("E: AI ai_network_create error - type=%d code=%d (%s)\r\n", err.type, err.code, yf_b200_last_error_text());
    return -1;
  }
  const ai_network_params params = AI_NETWORK_PARAMS_INIT(AI_NETWORK_DATA_WEIGHTS(ai_network_data_weights_get()),
                                                          AI_NETWORK_DATA_ACTIVATIONS(activations));
  if (!ai_network_init(network, &params)) {
    err = ai_network_get_error(network);
    printf("E: AI ai_network_init error - type=%d code=%d (%s)\r\n", err.type, err.code, yf_b200_last_error_text());
    return -1;
  }
  return 0;
}

/* one ai_network_run call per <= 65,535 images: ai_buffer.n_batches is 16-bit */
static int aiRun(const ai_i8* in_data, ai_i8* out_data, unsigned n) {
  ai_buffer ai_input[AI_NETWORK_IN_NUM] = AI_NETWORK_IN;
  ai_buffer ai_output[AI_NETWORK_OUT_NUM] = AI_NETWORK_OUT;
  while (n) {
    const unsigned nb = n > 65535u ? 65535u : n;
    ai_input[0].n_batches = (ai_u16)nb; ai_input[0].data = AI_HANDLE_PTR(in_data);
    ai_output[0].n_batches = (ai_u16)nb; ai_output[0].data = AI_HANDLE_PTR(out_data);
    if (ai_network_run(network, &ai_input[0], &ai_output[0]) != (ai_i32)nb) {
      ai_error err = ai_network_get_error(network);
      printf("E: AI ai_network_run error - type=%d code=%d (%s)\r\n", err.type, err.code, yf_b200_last_error_text());
      return -1;
    }
    in_data += (size_t)nb * AI_NETWORK_IN_1_SIZE; out_data += (size_t)nb * AI_NETWORK_OUT_1_SIZE; n -= nb;
  }
  return 0;
}

int main(int argc, char** argv) {
  const unsigned n = argc > 1 ? (unsigned)atoi(argv[1]) : 8;
  const float conf = argc > 2 ? (float)atof(argv[2]) : 0.7f, iou = argc > 3 ? (float)atof(argv[3]) : 0.4f;
  const unsigned max_det = 16;
  if (aiInit()) return 1;

  ai_u8* frames = (ai_u8*)malloc((size_t)n * 112 * 112 * 2);
  ai_i8* in_data = (ai_i8*)malloc((size_t)n * AI_NETWORK_IN_1_SIZE);
  ai_i8* out_data = (ai_i8*)malloc((size_t)n * AI_NETWORK_OUT_1_SIZE);
  yf_b200_det* dets = (yf_b200_det*)malloc(sizeof(yf_b200_det) * n * max_det);
  int32_t* counts = (int32_t*)malloc(sizeof(int32_t) * n);
  unsigned s = 12345;
  for (size_t i = 0; i < (size_t)n * 112 * 112 * 2; ++i) { s = s * 1103515245u + 12345u; frames[i] = (ai_u8)(s >> 16); }

  if (yf_b200_preprocess_rgb565(network, frames, in_data, n) < 0) { printf("E: preprocess: %s\r\n", yf_b200_last_error_text()); return 1; }
  if (aiRun(in_data, out_data, n)) return 1;
  if (yf_b200_decode(network, out_data, n, conf, iou, 0, dets, counts, max_det) < 0) { printf("E: decode: %s\r\n", yf_b200_last_error_text()); return 1; }

  for (unsigned f = 0; f < n; ++f) {                          /* main.c:46,53 + yoloface.c:143-148: clamp to the frame, double */
    printf("=== Frame %u ===\r\n----------------------------------------\r\n", f);
    for (int k = 0; k < counts[f]; ++k) {
      const yf_b200_det* d = &dets[(size_t)f * max_det + k];
      int c[4] = {(int)d->x1, (int)d->y1, (int)d->x2, (int)d->y2};
      for (int j = 0; j < 4; ++j) c[j] = (c[j] < 0 ? 0 : (c[j] > 55 ? 55 : c[j])) * 2;
      printf("[Face %d] BBox: [%d, %d, %d, %d], Conf: %.2f\r\n", k + 1, c[0], c[1], c[2], c[3], d->conf);
    }
    printf("----------------------------------------\r\n[INFO] Total faces detected: %d\r\n", counts[f]);
  }
  yf_b200_stats st;
  yf_b200_get_stats(network, &st);
  printf("[INFO] %llu images, %llu kernel launches on device %d (%d SMs), fused=%d\r\n", (unsigned long long)st.images,
         (unsigned long long)st.kernel_launches, st.device, st.sm_count, st.fused);
  ai_network_destroy(network);
  free(frames); free(in_data); free(out_data); free(dets); free(counts);
  return 0;
}
