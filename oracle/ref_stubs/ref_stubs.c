/* Host stand-ins for the board symbols the reference's yoloface.c links against: the camera frame
 * buffer, the face counter main.c owns, and an LCD rectangle call that records what was drawn.
 * Test infrastructure (oracle/_ref). */
#include <stdint.h>
#include <string.h>

uint8_t RGB_DATA[112 * 112 * 2];
uint8_t face_num;

#define YF_REF_MAX_RECTS 512
static int g_rects[YF_REF_MAX_RECTS][4];
static int g_nrects;

void LCD_DrawRectangle(uint16_t x1, uint16_t y1, uint16_t x2, uint16_t y2, uint16_t color) {
  (void)color;
  if (g_nrects < YF_REF_MAX_RECTS) {
    g_rects[g_nrects][0] = x1; g_rects[g_nrects][1] = y1; g_rects[g_nrects][2] = x2; g_rects[g_nrects][3] = y2;
    ++g_nrects;
  }
}
void yf_ref_reset(void) { g_nrects = 0; face_num = 0; }
int yf_ref_rects(int* dst, int cap) {
  int n = g_nrects < cap ? g_nrects : cap;
  memcpy(dst, g_rects, sizeof(int) * 4 * (size_t)n);
  return g_nrects;
}
