/* HYPREDRV_config.h -- build configuration of hypredrive_b200 (mirrors the generated header of
 * the reference build). */
#ifndef HYPREDRV_CONFIG_HEADER
#define HYPREDRV_CONFIG_HEADER
#define HYPREDRV_RELEASE_VERSION "0.2.0-b200"
#define HYPREDRV_B200_NATIVE 1
#endif
