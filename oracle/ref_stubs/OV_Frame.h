/* Stub of the OV2640 frame header. */
#ifndef YF_REF_STUB_OV_FRAME_H
#define YF_REF_STUB_OV_FRAME_H
#include <stdint.h>
#endif
