/* Stub: the camera frame the reference reads (RGB_DATA, 112x112 RGB565 byte pairs). */
#ifndef YF_REF_STUB_PICTURE_H
#define YF_REF_STUB_PICTURE_H
#include <stdint.h>
extern uint8_t RGB_DATA[112 * 112 * 2];
#endif
