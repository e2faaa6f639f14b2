#ifndef YF_REF_STUB_YOLOFACE_H
#define YF_REF_STUB_YOLOFACE_H
#include <stdint.h>
void resize_rgb565_uint8_112_to_56_direct(void);
void prepare_yolo_data(void);
void post_process(void);
int aiInit(void);
int aiRun(void);
#endif
