/* stand-in for <libavutil/avutil.h>; TEST INFRASTRUCTURE / build shim. */
#pragma once
#include <stdio.h>
#include <string.h>
#include "libavutil/frame.h"
