/* stand-in for <libavutil/avutil.h>; Build shim for hosts without FFmpeg headers. */
#pragma once
#include <stdio.h>
#include <string.h>
#include "libavutil/frame.h"
