/* stand-in for <libavutil/mathematics.h>; Build shim for hosts without FFmpeg headers. */
#pragma once
#include <math.h>
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
