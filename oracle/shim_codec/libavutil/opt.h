/* TEST INFRASTRUCTURE — stand-in for <libavutil/opt.h>; see libavcodec/avcodec.h here. */
#pragma once
int av_opt_set(void *obj, const char *name, const char *val, int search_flags);
