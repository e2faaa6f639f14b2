/* Embeds the int8 model the reference deploys (yoloface/tflite/yoloface_int8.tflite, kept as a
 * data asset under stm32h7-yolo_b200/assets/) into the shared library.  YF_MODEL_PATH is set by
 * the Makefile. */
__asm__(
    ".section .rodata\n"
    ".balign 16\n"
    ".global yf_embedded_model\n"
    "yf_embedded_model:\n"
    ".incbin \"" YF_MODEL_PATH "\"\n"
    ".global yf_embedded_model_end\n"
    "yf_embedded_model_end:\n"
    ".balign 4\n"
    ".global yf_embedded_model_len\n"
    "yf_embedded_model_len:\n"
    ".int yf_embedded_model_end - yf_embedded_model\n"
    ".text\n");
