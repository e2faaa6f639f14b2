/*
 * ai_platform.h -- B200 re-statement of the public X-CUBE-AI platform ABI used by the yoloface path.
 *
 * Layout- and value-compatible with stm32/Middlewares/ST/AI/Inc/ai_platform.h of the reference
 * (X-CUBE-AI 7.0.0) for exactly the items the hot path's callers touch:
 *   ai_buffer ............ ai_platform.h:517-525   (32 bytes on LP64; n_batches is 16-bit)
 *   ai_error ............. ai_platform.h:467-470   (8-bit type, 24-bit code, returned by value)
 *   ai_network_params .... ai_platform.h:348-357,606-608
 *   ai_network_report .... ai_platform.h:627-655
 *   buffer format codes .. ai_platform.h:261-270,392-413
 *   error type/code enums  ai_platform.h:546-586
 * A translation unit compiled against the reference's own header links against
 * libyoloface_b200.so unchanged (tests/test_dropin_reference_caller.py does exactly that).
 * Written from the ABI description; no text is shared with ST's header.
 */
#ifndef YF_B200_AI_PLATFORM_H
#define YF_B200_AI_PLATFORM_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
#define AI_API_DECLARE_BEGIN extern "C" {
#define AI_API_DECLARE_END }
#else
#define AI_API_DECLARE_BEGIN
#define AI_API_DECLARE_END
#endif

#define AI_API_ENTRY __attribute__((visibility("default")))
#define AI_ALIGNED(n) __attribute__((aligned(n)))
#define AI_DEPRECATED
#define AI_UNUSED(x) (void)(x);

/* ---- scalar aliases ------------------------------------------------------------------- */
typedef void* ai_handle;
typedef const void* ai_handle_const;
typedef float ai_float;
typedef double ai_double;
typedef bool ai_bool;
typedef char ai_char;
typedef uint32_t ai_size;
typedef uintptr_t ai_uptr;
typedef unsigned int ai_uint;
typedef uint8_t ai_u8;
typedef uint16_t ai_u16;
typedef uint32_t ai_u32;
typedef uint64_t ai_u64;
typedef int ai_int;
typedef int8_t ai_i8;
typedef int16_t ai_i16;
typedef int32_t ai_i32;
typedef int64_t ai_i64;
typedef uint32_t ai_signature;
typedef int32_t ai_buffer_format;

#define AI_HANDLE_PTR(p) ((ai_handle)(p))
#define AI_HANDLE_NULL AI_HANDLE_PTR(NULL)
#define AI_MAGIC_MARKER (0xA1FACADE)   /* brackets the weights table, network_data.c:395-401 */
#define AI_MAGIC_SIGNATURE (0xA1FACADE)
#define AI_FLAG_NONE (0x0)

/* ---- buffer format word: float[24] sign[23] type[17..20] bits[7..13] fbits+64[0..6] ---- */
#define AI_BUFFER_FMT_TYPE_NONE 0x0
#define AI_BUFFER_FMT_TYPE_FLOAT 0x1
#define AI_BUFFER_FMT_TYPE_Q 0x2
#define AI_BUFFER_FMT_TYPE_BOOL 0x3
#define AI_BUFFER_FMT_FLAG_CONST (0x1U << 30)
#define AI_BUFFER_FMT_FLAG_STATIC (0x1U << 29)
#define AI_BUFFER_FMT_FLAG_IS_IO (0x1U << 27)
#define AI_BUFFER_FMT_WORD(type, sign, flt, bits, fbits)                                       \
  ((ai_buffer_format)((((flt)&1) << 24) | (((sign)&1) << 23) | (((type)&0xF) << 17) |          \
                      (((bits)&0x7F) << 7) | (((fbits) + 64) & 0x7F)))
#define AI_BUFFER_FMT_GET(fmt) (((ai_buffer_format)(fmt)) & 0x01FFFFFF)
#define AI_BUFFER_FMT_GET_BITS(fmt) ((((ai_buffer_format)(fmt)) >> 7) & 0x7F)
#define AI_BUFFER_FMT_GET_SIGN(fmt) ((((ai_buffer_format)(fmt)) >> 23) & 0x1)
#define AI_BUFFER_FMT_GET_TYPE(fmt) ((((ai_buffer_format)(fmt)) >> 17) & 0xF)

enum {
  AI_BUFFER_FORMAT_NONE = AI_BUFFER_FMT_WORD(AI_BUFFER_FMT_TYPE_NONE, 0, 0, 0, 0),
  AI_BUFFER_FORMAT_FLOAT = AI_BUFFER_FMT_WORD(AI_BUFFER_FMT_TYPE_FLOAT, 1, 1, 32, 0),
  AI_BUFFER_FORMAT_U8 = AI_BUFFER_FMT_WORD(AI_BUFFER_FMT_TYPE_Q, 0, 0, 8, 0),
  AI_BUFFER_FORMAT_S8 = AI_BUFFER_FMT_WORD(AI_BUFFER_FMT_TYPE_Q, 1, 0, 8, 0),
  AI_BUFFER_FORMAT_S32 = AI_BUFFER_FMT_WORD(AI_BUFFER_FMT_TYPE_Q, 1, 0, 32, 0),
};

/* ---- descriptors ---------------------------------------------------------------------- */
typedef struct ai_error_ {
  ai_u32 type : 8;
  ai_u32 code : 24;
} ai_error;

typedef struct ai_intq_info_ {
  const ai_float* scale;
  ai_handle_const zeropoint;
} ai_intq_info;

typedef struct ai_intq_info_list_ {
  ai_u16 flags;
  ai_u16 size;
  const ai_intq_info* info;
} ai_intq_info_list;

#define AI_BUFFER_META_HAS_INTQ_INFO (0x1U << 0)
#define AI_BUFFER_META_FLAG_SCALE_FLOAT (0x1U << 0)
#define AI_BUFFER_META_FLAG_ZEROPOINT_U8 (0x1U << 1)
#define AI_BUFFER_META_FLAG_ZEROPOINT_S8 (0x1U << 2)

typedef struct ai_buffer_meta_info_ {
  ai_u32 flags;
  ai_intq_info_list* intq_info;
} ai_buffer_meta_info;

typedef struct ai_buffer_ {
  ai_buffer_format format;
  ai_u16 n_batches;   /* 16-bit: at most 65,535 images per ai_network_run call */
  ai_u16 height;
  ai_u16 width;
  ai_u32 channels;
  ai_handle data;
  ai_buffer_meta_info* meta_info;
} ai_buffer;

typedef struct ai_buffer_array_ {
  ai_u16 flags;
  ai_u16 size;
  ai_buffer* buffer;
} ai_buffer_array;

/* argument order: format, height, width, channels, n_batches, data (ai_platform.h:322-330) */
#define AI_BUFFER_OBJ_INIT(format_, h_, w_, ch_, n_batches_, data_)                            \
  { .format = (ai_buffer_format)(format_), .n_batches = (n_batches_), .height = (h_),          \
    .width = (w_), .channels = (ch_), .data = (ai_handle)(data_), .meta_info = NULL }
#define AI_BUFFER_SIZE(b) (((b)->width) * ((b)->height) * ((b)->channels))

typedef struct ai_network_params_ {
  union {
    struct { ai_buffer params; ai_buffer activations; };
    struct { ai_signature map_signature; ai_buffer_array map_weights; ai_buffer_array map_activations; };
  };
} ai_network_params;

#ifdef __cplusplus
#define AI_NETWORK_PARAMS_INIT(params_, activations_) { { { params_, activations_ } } }
#else
#define AI_NETWORK_PARAMS_INIT(params_, activations_) { .params = params_, .activations = activations_ }
#endif

typedef struct ai_platform_version_ {
  ai_u8 major, minor, micro, reserved;
} ai_platform_version;

typedef struct ai_network_report_ {
  const char* model_name;
  const char* model_signature;
  const char* model_datetime;
  const char* compile_datetime;
  const char* runtime_revision;
  ai_platform_version runtime_version;
  const char* tool_revision;
  ai_platform_version tool_version;
  ai_platform_version tool_api_version;
  ai_platform_version api_version;
  ai_platform_version interface_api_version;
  ai_u32 n_macc;
  ai_u16 n_inputs;
  ai_u16 n_outputs;
  ai_buffer* inputs;
  ai_buffer* outputs;
  union {
    struct { ai_buffer params; ai_buffer activations; };
    struct { ai_signature map_signature; ai_buffer_array map_weights; ai_buffer_array map_activations; };
  };
  ai_u32 n_nodes;
  ai_signature signature;
} ai_network_report;

/* ---- errors (first-error latch, cleared by ai_network_get_error) ----------------------- */
typedef enum {
  AI_ERROR_NONE = 0x00,
  AI_ERROR_TOOL_PLATFORM_API_MISMATCH = 0x01,
  AI_ERROR_TYPES_MISMATCH = 0x02,
  AI_ERROR_INVALID_HANDLE = 0x10,
  AI_ERROR_INVALID_STATE = 0x11,
  AI_ERROR_INVALID_INPUT = 0x12,
  AI_ERROR_INVALID_OUTPUT = 0x13,
  AI_ERROR_INVALID_PARAM = 0x14,
  AI_ERROR_INVALID_SIGNATURE = 0x15,
  AI_ERROR_INVALID_SIZE = 0x16,
  AI_ERROR_INVALID_VALUE = 0x17,
  AI_ERROR_INIT_FAILED = 0x30,
  AI_ERROR_ALLOCATION_FAILED = 0x31,
  AI_ERROR_DEALLOCATION_FAILED = 0x32,
  AI_ERROR_CREATE_FAILED = 0x33,
} ai_error_type;

typedef enum {
  AI_ERROR_CODE_NONE = 0x0000,
  AI_ERROR_CODE_NETWORK = 0x0010,
  AI_ERROR_CODE_NETWORK_PARAMS = 0x0011,
  AI_ERROR_CODE_NETWORK_WEIGHTS = 0x0012,
  AI_ERROR_CODE_NETWORK_ACTIVATIONS = 0x0013,
  AI_ERROR_CODE_LAYER = 0x0014,
  AI_ERROR_CODE_TENSOR = 0x0015,
  AI_ERROR_CODE_ARRAY = 0x0016,
  AI_ERROR_CODE_INVALID_PTR = 0x0017,
  AI_ERROR_CODE_INVALID_SIZE = 0x0018,
  AI_ERROR_CODE_INVALID_FORMAT = 0x0019,
  AI_ERROR_CODE_OUT_OF_RANGE = 0x0020,
  AI_ERROR_CODE_INVALID_BATCH = 0x0021,
  AI_ERROR_CODE_MISSED_INIT = 0x0030,
  AI_ERROR_CODE_IN_USE = 0x0040,
} ai_error_code;

#endif /* YF_B200_AI_PLATFORM_H */
