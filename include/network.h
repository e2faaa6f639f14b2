/*
 * network.h -- the drop-in boundary of the yoloface int8 hot path on B200.
 *
 * Same exported symbols, argument meaning and error behaviour as the X-CUBE-AI generated
 * stm32/X-CUBE-AI/App/network.h:103-213 of the reference; the implementation behind them is
 * hand-written sm_100a CUDA (libyoloface_b200.so) instead of ST's Cortex-M7 runtime.
 * Each entry point names the reference declaration it replaces.
 */
#ifndef YF_B200_NETWORK_H
#define YF_B200_NETWORK_H

#include "ai_platform.h"
#include "network_config.h"

#define AI_NETWORK_MODEL_NAME "network"                      /* network.h:29 */
#define AI_NETWORK_ORIGIN_MODEL_NAME "yoloface_int8"         /* network.h:30 */
#define AI_NETWORK_ACTIVATIONS_ALIGNMENT (4)

/* one int8 NHWC input [B,56,56,3] (scale 1/255, zp -128): network.h:38-52 */
#define AI_NETWORK_IN_NUM (1)
#define AI_NETWORK_IN_1_HEIGHT (56)
#define AI_NETWORK_IN_1_WIDTH (56)
#define AI_NETWORK_IN_1_CHANNEL (3)
#define AI_NETWORK_IN_1_SIZE (56 * 56 * 3)
#define AI_NETWORK_IN_1_SIZE_BYTES (AI_NETWORK_IN_1_SIZE * 1)
#define AI_NETWORK_IN { AI_BUFFER_OBJ_INIT(AI_BUFFER_FORMAT_S8, 56, 56, 3, 1, NULL), }
#define AI_NETWORK_IN_SIZE { AI_NETWORK_IN_1_SIZE, }
#define AI_NETWORK_IN_SIZE_BYTES { AI_NETWORK_IN_1_SIZE_BYTES, }

/* one int8 NHWC output [B,7,7,18] (scale 0.14218327403068542, zp -15): network.h:55-69 */
#define AI_NETWORK_OUT_NUM (1)
#define AI_NETWORK_OUT_1_HEIGHT (7)
#define AI_NETWORK_OUT_1_WIDTH (7)
#define AI_NETWORK_OUT_1_CHANNEL (18)
#define AI_NETWORK_OUT_1_SIZE (7 * 7 * 18)
#define AI_NETWORK_OUT_1_SIZE_BYTES (AI_NETWORK_OUT_1_SIZE * 1)
#define AI_NETWORK_OUT { AI_BUFFER_OBJ_INIT(AI_BUFFER_FORMAT_S8, 7, 7, 18, 1, NULL), }
#define AI_NETWORK_OUT_SIZE { AI_NETWORK_OUT_1_SIZE, }
#define AI_NETWORK_OUT_SIZE_BYTES { AI_NETWORK_OUT_1_SIZE_BYTES, }

#define AI_NETWORK_N_NODES (31)                              /* ST's fused c-node count, network.h:72 */

AI_API_DECLARE_BEGIN

/* replaces network.h:103-107 / network.c:3270 (deprecated alias of get_report) */
AI_DEPRECATED AI_API_ENTRY ai_bool ai_network_get_info(ai_handle network, ai_network_report* report);
/* replaces network.h:117-119 / network.c:3314 */
AI_API_ENTRY ai_bool ai_network_get_report(ai_handle network, ai_network_report* report);
/* replaces network.h:131-132 / network.c:3359: first error since the last call; reading clears it */
AI_API_ENTRY ai_error ai_network_get_error(ai_handle network);
/* replaces network.h:143-145 / network.c:3365.  network_config: NULL (defaults) or a yf_b200_config
 * wrapped in an ai_buffer (yoloface_b200.h).  ST's runtime hands out one static context
 * (g_network, network.c:36); here every call creates an independent context (one per GPU/config). */
AI_API_ENTRY ai_error ai_network_create(ai_handle* network, const ai_buffer* network_config);
/* replaces network.h:156-157 / network.c:3375: AI_HANDLE_NULL on success */
AI_API_ENTRY ai_handle ai_network_destroy(ai_handle network);
/* replaces network.h:172-174 / network.c:3381: params->params is the weights handle returned by
 * ai_network_data_weights_get() (marker table), or a raw pointer to the 11,304-byte blob */
AI_API_ENTRY ai_bool ai_network_init(ai_handle network, const ai_network_params* params);
/* replaces network.h:193-195 / network.c:3400: returns the number of batches processed, <=0 on
 * failure.  input/output .data may be host or device pointers. */
AI_API_ENTRY ai_i32 ai_network_run(ai_handle network, const ai_buffer* input, ai_buffer* output);
/* replaces network.h:209-211 / network.c:3407: run without reading the output back */
AI_API_ENTRY ai_i32 ai_network_forward(ai_handle network, const ai_buffer* input);

AI_API_DECLARE_END
#endif /* YF_B200_NETWORK_H */
