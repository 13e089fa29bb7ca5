/* unpaper_b200_types.h — ABI mirror of the value types that cross the
 * reference's backend boundary.
 *
 * The B200 backend is a drop-in for `const ImageBackend backend_cuda`
 * (reference imageprocess/backend.c:86-88, imageprocess/backend.h:19-57), whose
 * entry points take these structs BY VALUE.  Layout therefore has to agree,
 * field for field, with the reference headers cited beside each type.  When
 * this library is compiled inside the reference tree, define
 * UNPAPER_B200_WITH_REFERENCE_HEADERS and the reference's own headers are
 * used instead of this mirror.
 */
#pragma once

#ifdef UNPAPER_B200_WITH_REFERENCE_HEADERS
#include "imageprocess/backend.h"
#include "imageprocess/image.h"
#include "lib/options.h"
#else

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct AVFrame AVFrame;

/* imageprocess/primitives.h:12-15 */
typedef struct { int32_t x, y; } Point;
/* imageprocess/primitives.h:22-25 */
typedef struct { int32_t horizontal, vertical; } Delta;
/* imageprocess/primitives.h:39-42 */
typedef struct { bool horizontal, vertical; } Direction;
/* imageprocess/primitives.h:53-58 */
typedef struct { bool left, top, right, bottom; } Edges;
/* imageprocess/primitives.h:60-64 */
typedef struct { uint8_t r, g, b; } Pixel;
/* imageprocess/primitives.h:71-73: inclusive corners */
typedef struct { Point vertex[2]; } Rectangle;
/* imageprocess/primitives.h:75-78 */
typedef struct { int32_t width, height; } RectangleSize;
/* imageprocess/primitives.h:102-105 */
typedef struct { float x, y; } FloatPoint;

/* imageprocess/image.h:11-15 */
typedef struct {
  AVFrame *frame;
  Pixel background;
  uint8_t abs_black_threshold;
} Image;

/* imageprocess/interpolate.h:10-15 */
typedef enum {
  INTERP_NN,
  INTERP_LINEAR,
  INTERP_CUBIC,
  INTERP_FUNCTIONS_COUNT
} Interpolation;

/* imageprocess/blit.h:25-27 */
typedef int8_t RotationDirection;

/* constants.h:7-10 */
#define MAX_MASKS 100
#define MAX_POINTS 100
#define MAX_PAGES 2
#define MAX_WIPES MAX_MASKS

/* constants.h:18-23 */
typedef enum { LAYOUT_NONE, LAYOUT_SINGLE, LAYOUT_DOUBLE, LAYOUTS_COUNT } Layout;

/* imageprocess/filters.h:12-29 */
typedef struct {
  RectangleSize scan_size;
  Delta scan_step;
  struct { uint32_t horizontal, vertical; } scan_depth;
  Direction scan_direction;
  uint8_t abs_threshold;
  int32_t intensity;
  size_t exclusions_count;
  Rectangle *exclusions;
} BlackfilterParameters;

/* imageprocess/filters.h:41-47 */
typedef struct {
  RectangleSize scan_size;
  Delta scan_step;
  float intensity;
} BlurfilterParameters;

/* imageprocess/filters.h:58-64 */
typedef struct {
  RectangleSize scan_size;
  Delta scan_step;
  uint8_t abs_threshold;
} GrayfilterParameters;

/* imageprocess/deskew.h:13-20 */
typedef struct {
  float deskewScanRangeRad;
  float deskewScanStepRad;
  float deskewScanDeviationRad;
  int deskewScanSize;
  float deskewScanDepth;
  Edges scan_edges;
} DeskewParameters;

/* imageprocess/masks.h:14-37 */
typedef struct {
  RectangleSize scan_size;
  Delta scan_step;
  struct { int32_t horizontal, vertical; } scan_depth;
  Direction scan_direction;
  struct { float horizontal, vertical; } scan_threshold;
  int32_t minimum_width;
  int32_t maximum_width;
  int32_t minimum_height;
  int32_t maximum_height;
} MaskDetectionParameters;

/* imageprocess/masks.h:53-56 */
typedef struct {
  Edges alignment;
  Delta margin;
} MaskAlignmentParameters;

/* imageprocess/masks.h:67-70 */
typedef struct {
  size_t count;
  Rectangle areas[MAX_WIPES];
} Wipes;

/* imageprocess/masks.h:74-79 */
typedef struct { int32_t left, top, right, bottom; } Border;

/* imageprocess/masks.h:86-96 */
typedef struct {
  RectangleSize scan_size;
  Delta scan_step;
  struct { int32_t horizontal, vertical; } scan_threshold;
  Direction scan_direction;
} BorderScanParameters;

/* imageprocess/backend.h:19-57 — `name` + 20 entry points, in this order. */
typedef struct {
  const char *name;

  void (*wipe_rectangle)(Image image, Rectangle input_area, Pixel color);
  void (*copy_rectangle)(Image source, Image target, Rectangle source_area,
                         Point target_coords);
  void (*center_image)(Image source, Image target, Point target_origin,
                       RectangleSize target_size);
  void (*stretch_and_replace)(Image *pImage, RectangleSize size,
                              Interpolation interpolate_type);
  void (*resize_and_replace)(Image *pImage, RectangleSize size,
                             Interpolation interpolate_type);
  void (*flip_rotate_90)(Image *pImage, RotationDirection direction);
  void (*mirror)(Image image, Direction direction);
  void (*shift_image)(Image *pImage, Delta d);

  void (*apply_masks)(Image image, const Rectangle masks[], size_t masks_count,
                      Pixel color);
  void (*apply_wipes)(Image image, Wipes wipes, Pixel color);
  void (*apply_border)(Image image, const Border border, Pixel color);
  size_t (*detect_masks)(Image image, MaskDetectionParameters params,
                         const Point points[], size_t points_count,
                         Rectangle masks[]);
  void (*align_mask)(Image image, const Rectangle inside_area,
                     const Rectangle outside, MaskAlignmentParameters params);
  Border (*detect_border)(Image image, BorderScanParameters params,
                          const Rectangle outside_mask);

  void (*blackfilter)(Image image, BlackfilterParameters params);
  void (*blurfilter)(Image image, BlurfilterParameters params,
                     uint8_t abs_white_threshold);
  void (*noisefilter)(Image image, uint64_t intensity, uint8_t min_white_level);
  void (*grayfilter)(Image image, GrayfilterParameters params);

  float (*detect_rotation)(Image image, Rectangle mask,
                           const DeskewParameters params);
  void (*deskew)(Image source, Rectangle mask, float radians,
                 Interpolation interpolate_type);
} ImageBackend;

#ifdef __cplusplus
}
#endif

#endif /* UNPAPER_B200_WITH_REFERENCE_HEADERS */
