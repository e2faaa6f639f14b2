/* Stub of the board's LCD driver header so the reference's yoloface.c compiles on a host
 * (test infrastructure; see oracle/Makefile target `ref`). */
#ifndef YF_REF_STUB_LCD_H
#define YF_REF_STUB_LCD_H
#include <stdint.h>
#define RED 0xF800
void LCD_DrawRectangle(uint16_t x1, uint16_t y1, uint16_t x2, uint16_t y2, uint16_t color);
#endif
