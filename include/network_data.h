/*
 * network_data.h -- weights/activations descriptors of the drop-in boundary.
 * Replaces stm32/X-CUBE-AI/App/network_data.h:28-71 of the reference.
 */
#ifndef YF_B200_NETWORK_DATA_H
#define YF_B200_NETWORK_DATA_H

#include "ai_platform.h"
#include "network_config.h"

#define AI_NETWORK_DATA_CONFIG (NULL)
#define AI_NETWORK_DATA_ACTIVATIONS_SIZE (29784)   /* ST arena size; validated, not used: the  */
#define AI_NETWORK_DATA_ACTIVATIONS_COUNT (1)      /* B200 build keeps activations in HBM      */
#define AI_NETWORK_DATA_WEIGHTS_SIZE (11304)
#define AI_NETWORK_DATA_WEIGHTS_COUNT (1)

#define AI_NETWORK_DATA_ACTIVATIONS(ptr_)                                                      \
  AI_BUFFER_OBJ_INIT(AI_BUFFER_FORMAT_U8, AI_NETWORK_DATA_ACTIVATIONS_COUNT, 1,                \
                     AI_NETWORK_DATA_ACTIVATIONS_SIZE, 1, AI_HANDLE_PTR(ptr_))
#define AI_NETWORK_DATA_WEIGHTS(ptr_)                                                          \
  AI_BUFFER_OBJ_INIT(AI_BUFFER_FORMAT_U8 | AI_BUFFER_FMT_FLAG_CONST,                           \
                     AI_NETWORK_DATA_WEIGHTS_COUNT, 1, AI_NETWORK_DATA_WEIGHTS_SIZE, 1,        \
                     AI_HANDLE_PTR(ptr_))

AI_API_DECLARE_BEGIN
/* replaces network_data.h:60-62 / network_data.c:393-403: {MARKER, blob, MARKER} table whose blob
 * has ST's layout (offsets of network.c:3117-3263), regenerated from the embedded .tflite */
AI_DEPRECATED AI_API_ENTRY ai_handle ai_network_data_weights_get(void);
/* replaces network_data.h:70-71 / network_data.c:412-432 */
AI_API_ENTRY ai_bool ai_network_data_params_get(ai_handle network, ai_network_params* params);
AI_API_DECLARE_END
#endif
