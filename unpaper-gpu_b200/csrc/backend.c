/* backend.c — `backend_cuda`: the reference's 20-slot ImageBackend vtable
 * (imageprocess/backend.h:19-57) implemented on the B200 kernels.
 *
 * Signatures are exactly those of reference backend_cuda_internal.h:98-135
 * (images and parameter structs by value; stretch/resize/flip/shift replace
 * *pImage).  Every op makes the image device-resident, runs the same stage the
 * sheet engine runs — with a group of one page — and marks the device copy
 * newest.  Detectors read their small result record back once per call.
 * There is no CPU fallback: without a GPU the first call aborts
 * (rt.c: b200_rt_device).
 */
#define _GNU_SOURCE
#include <libavutil/frame.h>
#include <math.h>
#include <string.h>

#include "host.h"

/* ---- per-call context: one page, scratch from the cache --------------------- */

typedef struct {
  StageCtx sc;
  DPage hp;            /* host copy of the page record */
  DPage *dp;           /* device record */
  void *aux, *cls, *list, *u32, *stack, *jobs, *pre;
} Op;

static void op_begin(Op *o, Image *image, int need_aux, int need_nf, int u32_need, int need_stack) {
  memset(o, 0, sizeof(*o));
  image_ensure_cuda(image);
  DImg v;
  if (!b200_image_view(image, &v)) b200_fatal("unsupported pixel format %d", image->frame->format);
  ScratchNeed n;
  scratch_need_all(&n, v.w, v.h, v.fmt);
  o->hp.img = v;
  if (need_aux) {
    o->aux = b200_dev_alloc(n.aux_bytes);
    o->hp.aux = v;
    o->hp.aux.data = (uint8_t *)o->aux; o->hp.aux.pitch = n.aux_pitch; o->hp.aux.h = n.aux_h; o->hp.aux.w = v.w + 64;
  }
  if (need_nf) {
    o->cls = b200_dev_alloc(n.cls_bytes);
    o->list = b200_dev_alloc((size_t)need_nf * 4);
    o->hp.cls = (uint8_t *)o->cls; o->hp.list = (uint32_t *)o->list; o->hp.list_cap = need_nf;
  }
  if (u32_need > 0) {
    o->u32 = b200_dev_alloc((size_t)u32_need * 4);
    o->hp.u32 = (uint32_t *)o->u32; o->hp.u32_cap = u32_need;
  }
  if (need_stack) {
    o->stack = b200_dev_alloc((size_t)n.stack_cap * 32);
    o->hp.stack = (uint64_t *)o->stack; o->hp.stack_cap = n.stack_cap;
  }
  size_t jb = 2 * sizeof(DFillJob) + 2 * sizeof(DCopyJob) + sizeof(DMaskJob) + sizeof(DPage) + 256;
  o->jobs = b200_dev_alloc(jb);
  uint8_t *jp = (uint8_t *)o->jobs;
  o->dp = (DPage *)jp; jp += (sizeof(DPage) + 63) & ~(size_t)63;
  o->sc.fillA = (DFillJob *)jp; jp += sizeof(DFillJob);
  o->sc.fillB = (DFillJob *)jp; jp += sizeof(DFillJob);
  o->sc.copyA = (DCopyJob *)jp; jp += sizeof(DCopyJob);
  o->sc.copyB = (DCopyJob *)jp; jp += sizeof(DCopyJob);
  o->sc.maskJ = (DMaskJob *)jp;
  o->sc.st = b200_rt_stream();
  o->sc.npages = 1;
  o->sc.pages = o->dp;
  o->sc.w = v.w; o->sc.h = v.h; o->sc.fmt = v.fmt;
  o->sc.rows_aligned16 = ((v.pitch & 15) == 0) && (((uintptr_t)v.data & 15) == 0);
  o->hp.rot_cos[0] = 1.0f;
}

static void op_push(Op *o) {   /* host page record -> device */
  CUDA_OK(cudaMemcpyAsync(o->dp, &o->hp, sizeof(DPage), cudaMemcpyHostToDevice, o->sc.st));
}
static void op_pull(Op *o) {   /* device page record -> host (synchronises) */
  CUDA_OK(cudaMemcpyAsync(&o->hp, o->dp, sizeof(DPage), cudaMemcpyDeviceToHost, o->sc.st));
  CUDA_OK(cudaStreamSynchronize(o->sc.st));
  CUDA_OK(cudaGetLastError());
}
static void op_end(Op *o, Image *image, bool modified) {
  CUDA_OK(cudaStreamSynchronize(o->sc.st));
  CUDA_OK(cudaGetLastError());
  if (modified) image_mark_cuda_dirty(image);
  if (o->aux) b200_dev_free(o->aux);
  if (o->cls) b200_dev_free(o->cls);
  if (o->list) b200_dev_free(o->list);
  if (o->u32) b200_dev_free(o->u32);
  if (o->stack) b200_dev_free(o->stack);
  if (o->jobs) b200_dev_free(o->jobs);
  if (o->pre) b200_dev_free(o->pre);
}
static void op_check(const Op *o, const char *what) {
  if (o->hp.error) b200_fatal("%s: device-side failure flags 0x%x", what, o->hp.error);
}

static DRect drect(Rectangle r) { return (DRect){r.vertex[0].x, r.vertex[0].y, r.vertex[1].x, r.vertex[1].y}; }
static Rectangle urect(DRect r) { return (Rectangle){{{r.x0, r.y0}, {r.x1, r.y1}}}; }
static DRect drect_norm(Rectangle r) {
  DRect d = drect(r);
  DRect n = {d.x0 < d.x1 ? d.x0 : d.x1, d.y0 < d.y1 ? d.y0 : d.y1, d.x0 > d.x1 ? d.x0 : d.x1, d.y0 > d.y1 ? d.y0 : d.y1};
  return n;
}

/* small helper: run explicit blit jobs given on the host */
static void push_fill(Op *o, DFillJob *slot, DImg img, DRect r, Pixel c) {
  DFillJob j; memset(&j, 0, sizeof(j));
  j.img = img; j.r = r; j.c[0] = c.r; j.c[1] = c.g; j.c[2] = c.b; j.enabled = 1;
  CUDA_OK(cudaMemcpyAsync(slot, &j, sizeof(j), cudaMemcpyHostToDevice, o->sc.st));
}
static void push_copy(Op *o, DCopyJob *slot, DImg src, DImg dst, DRect area, int tx, int ty) {
  DCopyJob j; memset(&j, 0, sizeof(j));
  j.src = src; j.dst = dst; j.area = area; j.tx = tx; j.ty = ty; j.enabled = 1;
  CUDA_OK(cudaMemcpyAsync(slot, &j, sizeof(j), cudaMemcpyHostToDevice, o->sc.st));
}
static int bpp_of(int fmt) { return fmt == DF_GRAY8 ? 1 : fmt == DF_Y400A ? 2 : fmt == DF_RGB24 ? 3 : 1; }

/* ---- blit ops ---------------------------------------------------------------- */

/* wipe_rectangle (blit.c:20-24).  A wipe that covers the whole image makes the
 * previous contents irrelevant, so an unallocated device copy is created
 * WITHOUT uploading the (possibly uninitialised) host buffer. */
static void wipe_rectangle_b200(Image image, Rectangle input_area, Pixel color) {
  if (!image.frame) return;
  DRect r = drect_norm(input_area);
  bool whole = r.x0 <= 0 && r.y0 <= 0 && r.x1 >= image.frame->width - 1 && r.y1 >= image.frame->height - 1;
  if (whole) { image_ensure_cuda_alloc(&image); image_mark_cuda_dirty(&image); }
  Op o; op_begin(&o, &image, 0, 0, 0, 0);
  push_fill(&o, o.sc.fillA, o.hp.img, r, color);
  b200k_fill_jobs(o.sc.st, o.sc.fillA, 1, o.hp.img.w, o.hp.img.h);
  op_end(&o, &image, true);
}

static void copy_rectangle_b200(Image source, Image target, Rectangle source_area, Point target_coords) {
  if (!source.frame || !target.frame) return;
  image_ensure_cuda(&source);
  Op o; op_begin(&o, &target, 0, 0, 0, 0);
  DImg sv;
  if (!b200_image_view(&source, &sv)) b200_fatal("unsupported source pixel format");
  push_copy(&o, o.sc.copyA, sv, o.hp.img, drect(source_area), target_coords.x, target_coords.y);
  b200k_copy_jobs(o.sc.st, o.sc.copyA, 1, sv.w * bpp_of(sv.fmt), sv.h);
  op_end(&o, &target, true);
}

/* center_image (blit.c:175-202) */
static void center_image_b200(Image source, Image target, Point target_origin, RectangleSize target_size) {
  if (!source.frame || !target.frame) return;
  Point source_origin = {0, 0};
  RectangleSize source_size = {source.frame->width, source.frame->height};
  if (source_size.width < target_size.width || source_size.height < target_size.height) {
    Rectangle r = {{target_origin, {target_origin.x + target_size.width - 1, target_origin.y + target_size.height - 1}}};
    wipe_rectangle_b200(target, r, target.background);
  }
  if (source_size.width <= target_size.width) target_origin.x += (target_size.width - source_size.width) / 2;
  else { source_origin.x += (source_size.width - target_size.width) / 2; source_size.width = target_size.width; }
  if (source_size.height <= target_size.height) target_origin.y += (target_size.height - source_size.height) / 2;
  else { source_origin.y += (source_size.height - target_size.height) / 2; source_size.height = target_size.height; }
  Rectangle area = {{source_origin, {source_origin.x + source_size.width - 1, source_origin.y + source_size.height - 1}}};
  copy_rectangle_b200(source, target, area, target_origin);
}

/* a fresh device-resident image compatible with `src` (create_compatible_image, image.c:56-59) */
static Image new_compatible(Image src, RectangleSize size, bool fill) {
  Image img = {.frame = av_frame_alloc(), .background = src.background, .abs_black_threshold = src.abs_black_threshold};
  img.frame->width = size.width; img.frame->height = size.height; img.frame->format = src.frame->format;
  if (av_frame_get_buffer(img.frame, 8) < 0) b200_fatal("unable to allocate image buffer");
  image_ensure_cuda_alloc(&img);
  image_mark_cuda_dirty(&img);
  if (fill) {
    Rectangle full = {{{0, 0}, {size.width - 1, size.height - 1}}};
    wipe_rectangle_b200(img, full, img.background);
  } else {
    /* defined contents for formats whose writers touch only part of a byte */
    DImg v; b200_image_view(&img, &v);
    CUDA_OK(cudaMemsetAsync(v.data, 0, (size_t)v.pitch * v.h, b200_rt_stream()));
  }
  return img;
}
static void replace_with(Image *pImage, Image *fresh) {
  image_cuda_release(pImage);
  av_frame_free(&pImage->frame);
  *pImage = *fresh;
  fresh->frame = NULL;
}

/* compare_sizes (primitives.c:70-80) == 0 */
static bool same_size(Image img, RectangleSize s) { return img.frame->width == s.width && img.frame->height == s.height; }

static void stretch_and_replace_b200(Image *pImage, RectangleSize size, Interpolation interp) {
  if (!pImage || !pImage->frame) return;
  if (same_size(*pImage, size)) return;   /* blit.c:232-233 */
  image_ensure_cuda(pImage);
  Image target = new_compatible(*pImage, size, false);
  DImg s, d;
  b200_image_view(pImage, &s); b200_image_view(&target, &d);
  float hr = (float)s.w / (float)d.w, vr = (float)s.h / (float)d.h;   /* blit.c:213-216 */
  cudaStream_t st = b200_rt_stream();
  b200k_stretch(st, s, d, hr, vr, (int)interp);
  CUDA_OK(cudaStreamSynchronize(st));
  replace_with(pImage, &target);
}

static void resize_and_replace_b200(Image *pImage, RectangleSize size, Interpolation interp) {
  if (!pImage || !pImage->frame) return;
  RectangleSize image_size = {pImage->frame->width, pImage->frame->height};
  if (same_size(*pImage, size)) return;
  /* blit.c:255-270 */
  const float horizontal_ratio = (float)size.width / (float)image_size.width;
  const float vertical_ratio = (float)size.height / (float)image_size.height;
  RectangleSize stretch_size;
  if (horizontal_ratio < vertical_ratio) stretch_size = (RectangleSize){size.width, image_size.height * horizontal_ratio};
  else if (vertical_ratio < horizontal_ratio) stretch_size = (RectangleSize){image_size.width * vertical_ratio, size.height};
  else stretch_size = size;
  stretch_and_replace_b200(pImage, stretch_size, interp);
  if (size.width == stretch_size.width && size.height == stretch_size.height) return;
  Image resized = new_compatible(*pImage, size, true);
  center_image_b200(*pImage, resized, (Point){0, 0}, size);
  replace_with(pImage, &resized);
}

static void flip_rotate_90_b200(Image *pImage, RotationDirection direction) {
  if (!pImage || !pImage->frame) return;
  image_ensure_cuda(pImage);
  RectangleSize ns = {pImage->frame->height, pImage->frame->width};
  Image target = new_compatible(*pImage, ns, false);
  DImg s, d;
  b200_image_view(pImage, &s); b200_image_view(&target, &d);
  cudaStream_t st = b200_rt_stream();
  b200k_rotate90(st, s, d, (int)direction);
  CUDA_OK(cudaStreamSynchronize(st));
  replace_with(pImage, &target);
}

static void mirror_b200(Image image, Direction direction) {
  if (!image.frame || (!direction.horizontal && !direction.vertical)) return;
  Op o; op_begin(&o, &image, 0, 0, 0, 0);
  b200k_mirror(o.sc.st, o.hp.img, direction.horizontal, direction.vertical);
  op_end(&o, &image, true);
}

static void shift_image_b200(Image *pImage, Delta d) {   /* blit.c:360-368 */
  if (!pImage || !pImage->frame) return;
  RectangleSize sz = {pImage->frame->width, pImage->frame->height};
  Image fresh = new_compatible(*pImage, sz, true);
  Rectangle full = {{{0, 0}, {sz.width - 1, sz.height - 1}}};
  copy_rectangle_b200(*pImage, fresh, full, (Point){d.horizontal, d.vertical});
  replace_with(pImage, &fresh);
}

/* ---- masks / wipes / borders ----------------------------------------------- */

static void apply_masks_b200(Image image, const Rectangle masks[], size_t masks_count, Pixel color) {
  if (!image.frame || masks_count == 0) return;   /* masks.c:313-315 */
  Op o; op_begin(&o, &image, 0, 0, 0, 0);
  DRect *h = (DRect *)malloc(masks_count * sizeof(DRect));
  for (size_t i = 0; i < masks_count; i++) h[i] = drect(masks[i]);
  DRect *d = (DRect *)blob_upload(h, masks_count * sizeof(DRect));
  free(h);
  DMaskJob j; memset(&j, 0, sizeof(j));
  j.img = o.hp.img; j.rects = d; j.nrects = (int)masks_count; j.c[0] = color.r; j.c[1] = color.g; j.c[2] = color.b; j.enabled = 1;
  CUDA_OK(cudaMemcpyAsync(o.sc.maskJ, &j, sizeof(j), cudaMemcpyHostToDevice, o.sc.st));
  b200k_apply_masks(o.sc.st, o.sc.maskJ, 1, o.hp.img.w, o.hp.img.h);
  op_end(&o, &image, true);
  b200_dev_free(d);
}

/* apply_wipes (masks.c:337-345): rectangles are NOT normalised and NOT clipped
 * (set_pixel drops what falls outside) */
static void apply_wipes_b200(Image image, Wipes wipes, Pixel color) {
  if (!image.frame || wipes.count == 0) return;
  Op o; op_begin(&o, &image, 0, 0, 0, 0);
  /* all rectangles in one launch (same colour, so their order does not matter) */
  DFillJob *h = (DFillJob *)calloc(wipes.count, sizeof(DFillJob));
  for (size_t i = 0; i < wipes.count; i++) {
    h[i].img = o.hp.img; h[i].r = drect(wipes.areas[i]);
    h[i].c[0] = color.r; h[i].c[1] = color.g; h[i].c[2] = color.b; h[i].enabled = 1;
  }
  DFillJob *d = (DFillJob *)blob_upload(h, wipes.count * sizeof(DFillJob));
  free(h);
  b200k_fill_jobs(o.sc.st, d, (int)wipes.count, o.hp.img.w, o.hp.img.h);
  op_end(&o, &image, true);
  b200_dev_free(d);
}

static void apply_border_b200(Image image, const Border border, Pixel color) {   /* masks.c:370-382 */
  if (!image.frame) return;
  if (border.left == 0 && border.top == 0 && border.right == 0 && border.bottom == 0) return;
  Rectangle mask = {{{border.left, border.top},
                     {image.frame->width - border.right - 1, image.frame->height - border.bottom - 1}}};
  apply_masks_b200(image, &mask, 1, color);
}

static size_t detect_masks_b200(Image image, MaskDetectionParameters params, const Point points[],
                                size_t points_count, Rectangle masks[]) {
  if (!image.frame) return 0;
  if (!params.scan_direction.horizontal && !params.scan_direction.vertical) return 0;   /* masks.c:182-184 */
  size_t total = 0;
  /* the device record holds D_MAX_MASKS points; larger lists go in slices */
  for (size_t base = 0; base < points_count; base += D_MAX_MASKS) {
    int n = (int)(points_count - base < D_MAX_MASKS ? points_count - base : D_MAX_MASKS);
    MaskPlan pl;
    if (mask_plan_build(&pl, image.frame->width, image.frame->height, &params, points + base, n)) b200_fatal("%s", unpaper_b200_last_error());
    Op o; op_begin(&o, &image, 0, 0, pl.u32_need, 0);
    o.hp.point_count = n;
    for (int i = 0; i < n; i++) { o.hp.px[i] = points[base + i].x; o.hp.py[i] = points[base + i].y; }
    op_push(&o);
    stage_detect_masks(&o.sc, &pl);
    op_pull(&o);
    op_check(&o, "detect_masks");
    for (int i = 0; i < n; i++) masks[base + i] = urect(o.hp.masks[i]);
    total += (size_t)o.hp.mask_count;
    op_end(&o, &image, false);
    mask_plan_free(&pl);
  }
  return total;
}

/* shared by center_mask-like moves: copy area -> temp, wipe area, paste at target */
static void move_area(Image image, Rectangle area, Point target) {
  Op o; op_begin(&o, &image, 1, 0, 0, 0);
  int w = abs(area.vertex[0].x - area.vertex[1].x) + 1, h = abs(area.vertex[0].y - area.vertex[1].y) + 1;
  DImg aux = o.hp.aux;
  int bpp = aux.fmt == DF_GRAY8 ? 1 : aux.fmt == DF_Y400A ? 2 : aux.fmt == DF_RGB24 ? 3 : 0;
  int pitch = bpp ? ((w * bpp + 15) & ~15) : (((w + 7) / 8 + 15) & ~15);
  if ((long long)pitch * h > (long long)aux.pitch * aux.h) b200_fatal("move_area: area larger than the image + 64 px");
  aux.w = w; aux.h = h; aux.pitch = pitch;
  Pixel bg = image.background;
  push_fill(&o, o.sc.fillA, aux, (DRect){0, 0, w - 1, h - 1}, bg);
  push_copy(&o, o.sc.copyA, o.hp.img, aux, drect(area), 0, 0);
  push_fill(&o, o.sc.fillB, o.hp.img, drect_norm(area), bg);
  push_copy(&o, o.sc.copyB, aux, o.hp.img, (DRect){0, 0, w - 1, h - 1}, target.x, target.y);
  int aw = o.hp.img.w + 64, ah = o.hp.img.h + 64;
  b200k_fill_jobs(o.sc.st, o.sc.fillA, 1, aw, ah);
  b200k_copy_jobs(o.sc.st, o.sc.copyA, 1, aw * bpp_of(aux.fmt), ah);
  b200k_fill_jobs(o.sc.st, o.sc.fillB, 1, aw, ah);
  b200k_copy_jobs(o.sc.st, o.sc.copyB, 1, aw * bpp_of(aux.fmt), ah);
  op_end(&o, &image, true);
}

static void align_mask_b200(Image image, const Rectangle inside_area, const Rectangle outside,
                            MaskAlignmentParameters params) {   /* masks.c:265-300 */
  if (!image.frame) return;
  int w = abs(inside_area.vertex[0].x - inside_area.vertex[1].x) + 1;
  int h = abs(inside_area.vertex[0].y - inside_area.vertex[1].y) + 1;
  Point target;
  if (params.alignment.left) target.x = outside.vertex[0].x + params.margin.horizontal;
  else if (params.alignment.right) target.x = outside.vertex[1].x - w - params.margin.horizontal;
  else target.x = (outside.vertex[0].x + outside.vertex[1].x - w) / 2;
  if (params.alignment.top) target.y = outside.vertex[0].y + params.margin.vertical;
  else if (params.alignment.bottom) target.y = outside.vertex[1].y - h - params.margin.vertical;
  else target.y = (outside.vertex[0].y + outside.vertex[1].y - h) / 2;
  move_area(image, inside_area, target);
}

/* center_mask is not a vtable slot in the reference (masks.c:222-249 composes
 * copy/wipe/copy); exported for the host API and the engine tests. */
void b200_center_mask(Image image, const Point center, const Rectangle area) {
  if (!image.frame) return;
  int w = abs(area.vertex[0].x - area.vertex[1].x) + 1, h = abs(area.vertex[0].y - area.vertex[1].y) + 1;
  Point target = {center.x + (-w / 2), center.y + (-h / 2)};
  int W = image.frame->width, H = image.frame->height;
  bool inside = target.x >= 0 && target.y >= 0 && target.x < W && target.y < H &&
                target.x + w - 1 >= 0 && target.x + w - 1 < W && target.y + h - 1 >= 0 && target.y + h - 1 < H;
  if (!inside) return;
  move_area(image, area, target);
}

static Border detect_border_b200(Image image, BorderScanParameters params, const Rectangle outside_mask) {
  Border b = {0, 0, 0, 0};
  if (!image.frame) return b;
  BorderPlan pl;
  if (border_plan_build(&pl, image.frame->width, image.frame->height, &params, &outside_mask, 1, image.abs_black_threshold))
    b200_fatal("%s", unpaper_b200_last_error());
  Op o; op_begin(&o, &image, 0, 0, pl.u32_need, 0);
  o.hp.outside_count = 1;
  o.hp.outside[0] = drect(outside_mask);
  op_push(&o);
  stage_detect_border(&o.sc, &pl);
  op_pull(&o);
  b = (Border){o.hp.border[0].left, o.hp.border[0].top, o.hp.border[0].right, o.hp.border[0].bottom};
  op_end(&o, &image, false);
  border_plan_free(&pl);
  return b;
}

/* ---- filters ------------------------------------------------------------------ */

static void blackfilter_b200(Image image, BlackfilterParameters params) {
  if (!image.frame) return;
  BfPlan pl;
  if (bf_plan_build(&pl, image.frame->width, image.frame->height, &params, image.abs_black_threshold))
    b200_fatal("%s", unpaper_b200_last_error());
  Op o; op_begin(&o, &image, 0, 0, pl.u32_need, 1);
  op_push(&o);
  stage_blackfilter(&o.sc, &pl);
  op_pull(&o);
  op_check(&o, "blackfilter");
  op_end(&o, &image, true);
  bf_plan_free(&pl);
}

static void blurfilter_b200(Image image, BlurfilterParameters params, uint8_t abs_white_threshold) {
  if (!image.frame) return;
  BlurPlan pl;
  if (blur_plan_build(&pl, image.frame->width, image.frame->height, &params, abs_white_threshold))
    b200_fatal("%s", unpaper_b200_last_error());
  Op o; op_begin(&o, &image, 0, 0, pl.u32_need, 0);
  op_push(&o);
  stage_blurfilter(&o.sc, &pl);
  op_end(&o, &image, true);
  blur_plan_free(&pl);
}

static void noisefilter_b200(Image image, uint64_t intensity, uint8_t min_white_level) {
  if (!image.frame) return;
  ScratchNeed n;
  scratch_need_all(&n, image.frame->width, image.frame->height, 0);
  n.list_cap = nf_list_cap(image.frame->width, image.frame->height, intensity);
  Op o; op_begin(&o, &image, 0, n.list_cap, 0, 0);
  op_push(&o);
  if (stage_noisefilter(&o.sc, intensity, min_white_level)) b200_fatal("%s", unpaper_b200_last_error());
  op_pull(&o);
  if (o.hp.error & DERR_LIST_OVERFLOW)
    b200_fatal("noisefilter: mutable-pixel list overflow (%u > %d)", o.hp.list_n, n.list_cap);
  op_check(&o, "noisefilter");
  op_end(&o, &image, true);
}

static void grayfilter_b200(Image image, GrayfilterParameters params) {
  if (!image.frame) return;
  GrayPlan pl;
  if (gray_plan_build(&pl, image.frame->width, image.frame->height, &params, image.abs_black_threshold))
    b200_fatal("%s", unpaper_b200_last_error());
  Op o; op_begin(&o, &image, 0, 0, pl.u32_need, 0);
  op_push(&o);
  if (stage_grayfilter(&o.sc, &pl)) b200_fatal("%s", unpaper_b200_last_error());
  op_end(&o, &image, true);
}

/* ---- deskew -------------------------------------------------------------------- */

static float detect_rotation_b200(Image image, Rectangle mask, const DeskewParameters params) {
  if (!image.frame) return 0.0f;
  RotPlan pl;
  if (rot_plan_build(&pl, image.frame->width, image.frame->height, &params, 1, false))
    b200_fatal("%s", unpaper_b200_last_error());
  Op o; op_begin(&o, &image, 0, 0, pl.u32_need, 0);
  o.pre = b200_dev_alloc((size_t)pl.pre_need * 4);
  o.hp.pre = (uint32_t *)o.pre; o.hp.pre_cap = pl.pre_need;
  o.hp.mask_count = 1;
  o.hp.masks[0] = drect(mask);
  op_push(&o);
  if (stage_detect_rotation(&o.sc, &pl, 1)) b200_fatal("%s", unpaper_b200_last_error());
  op_pull(&o);
  /* the float tail runs on the host with the host's libm, like the reference
   * (deskew.c:218-240; backend_cuda_deskew.c keeps it on the host too) */
  float r = rot_finalize_host(&pl, o.hp.rot_angle_idx[0]);
  op_end(&o, &image, false);
  rot_plan_free(&pl);
  return r;
}

static void deskew_b200(Image source, Rectangle mask, float radians, Interpolation interp) {
  if (!source.frame) return;
  Op o; op_begin(&o, &source, 1, 0, 0, 0);
  int ink_cells = ((o.hp.img.w + D_INK_CELL - 1) / D_INK_CELL) * ((o.hp.img.h + D_INK_CELL - 1) / D_INK_CELL);
  o.pre = b200_dev_alloc((size_t)ink_cells + 64);   /* freed with the other scratch */
  o.hp.ink = (uint8_t *)o.pre; o.hp.ink_cap = ink_cells;
  o.hp.mask_count = 1;
  o.hp.masks[0] = drect(mask);
  o.hp.rotation[0] = radians;
  o.hp.rot_sin[0] = sinf(-radians);   /* deskew.c:260-261 with rotate(..., -radians, ...) */
  o.hp.rot_cos[0] = cosf(-radians);
  o.hp.rot_apply[0] = 1;
  op_push(&o);
  stage_deskew(&o.sc, (int)interp, 1);
  op_pull(&o);
  op_check(&o, "deskew");
  op_end(&o, &source, true);
}

const ImageBackend backend_cuda = {
    .name = "cuda",
    .wipe_rectangle = wipe_rectangle_b200,
    .copy_rectangle = copy_rectangle_b200,
    .center_image = center_image_b200,
    .stretch_and_replace = stretch_and_replace_b200,
    .resize_and_replace = resize_and_replace_b200,
    .flip_rotate_90 = flip_rotate_90_b200,
    .mirror = mirror_b200,
    .shift_image = shift_image_b200,
    .apply_masks = apply_masks_b200,
    .apply_wipes = apply_wipes_b200,
    .apply_border = apply_border_b200,
    .detect_masks = detect_masks_b200,
    .align_mask = align_mask_b200,
    .detect_border = detect_border_b200,
    .blackfilter = blackfilter_b200,
    .blurfilter = blurfilter_b200,
    .noisefilter = noisefilter_b200,
    .grayfilter = grayfilter_b200,
    .detect_rotation = detect_rotation_b200,
    .deskew = deskew_b200,
};
