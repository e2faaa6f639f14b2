/* network_config.h -- replaces stm32/X-CUBE-AI/App/network_config.h:28-31 (tool/version macros). */
#ifndef YF_B200_NETWORK_CONFIG_H
#define YF_B200_NETWORK_CONFIG_H
#define AI_TOOLS_VERSION_MAJOR 7
#define AI_TOOLS_VERSION_MINOR 0
#define AI_TOOLS_VERSION_MICRO 0
#define AI_TOOLS_API_VERSION_MAJOR 1
#define AI_TOOLS_API_VERSION_MINOR 4
#define AI_TOOLS_API_VERSION_MICRO 0
#define YF_B200_BACKEND "sm_100a"
#endif
