"""ctypes mirror of include/unpaper_b200.h / unpaper_b200_types.h.

``HostOps`` binds the 21 host-buffer entry points of any library that exports
them under a common prefix (``unpaper_b200_host_*`` for the CUDA product; the
test suite binds its CPU checkers through the same class).
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

FMT_RGB24, FMT_GRAY8, FMT_MONOWHITE, FMT_MONOBLACK, FMT_Y400A = 2, 8, 9, 10, 58
INTERP_NN, INTERP_LINEAR, INTERP_CUBIC = 0, 1, 2
LAYOUT_NONE, LAYOUT_SINGLE, LAYOUT_DOUBLE = 0, 1, 2
MAX_MASKS = 100
MAX_PAGES = 2
TRACE_MAX_MASKS = 8


def _S(name, fields):
    return type(name, (C.Structure,), {"_fields_": fields})


Point = _S("Point", [("x", C.c_int32), ("y", C.c_int32)])
Delta = _S("Delta", [("horizontal", C.c_int32), ("vertical", C.c_int32)])
Direction = _S("Direction", [("horizontal", C.c_bool), ("vertical", C.c_bool)])
Edges = _S("Edges", [("left", C.c_bool), ("top", C.c_bool), ("right", C.c_bool), ("bottom", C.c_bool)])
Pixel = _S("Pixel", [("r", C.c_uint8), ("g", C.c_uint8), ("b", C.c_uint8)])
Rectangle = _S("Rectangle", [("vertex", Point * 2)])
RectangleSize = _S("RectangleSize", [("width", C.c_int32), ("height", C.c_int32)])
_U2 = _S("_U2", [("horizontal", C.c_uint32), ("vertical", C.c_uint32)])
_I2 = _S("_I2", [("horizontal", C.c_int32), ("vertical", C.c_int32)])
_F2 = _S("_F2", [("horizontal", C.c_float), ("vertical", C.c_float)])

BlackfilterParameters = _S("BlackfilterParameters", [
    ("scan_size", RectangleSize), ("scan_step", Delta), ("scan_depth", _U2),
    ("scan_direction", Direction), ("abs_threshold", C.c_uint8), ("intensity", C.c_int32),
    ("exclusions_count", C.c_size_t), ("exclusions", C.POINTER(Rectangle))])
BlurfilterParameters = _S("BlurfilterParameters", [
    ("scan_size", RectangleSize), ("scan_step", Delta), ("intensity", C.c_float)])
GrayfilterParameters = _S("GrayfilterParameters", [
    ("scan_size", RectangleSize), ("scan_step", Delta), ("abs_threshold", C.c_uint8)])
DeskewParameters = _S("DeskewParameters", [
    ("deskewScanRangeRad", C.c_float), ("deskewScanStepRad", C.c_float),
    ("deskewScanDeviationRad", C.c_float), ("deskewScanSize", C.c_int),
    ("deskewScanDepth", C.c_float), ("scan_edges", Edges)])
MaskDetectionParameters = _S("MaskDetectionParameters", [
    ("scan_size", RectangleSize), ("scan_step", Delta), ("scan_depth", _I2),
    ("scan_direction", Direction), ("scan_threshold", _F2),
    ("minimum_width", C.c_int32), ("maximum_width", C.c_int32),
    ("minimum_height", C.c_int32), ("maximum_height", C.c_int32)])
MaskAlignmentParameters = _S("MaskAlignmentParameters", [("alignment", Edges), ("margin", Delta)])
Wipes = _S("Wipes", [("count", C.c_size_t), ("areas", Rectangle * MAX_MASKS)])
Border = _S("Border", [("left", C.c_int32), ("top", C.c_int32), ("right", C.c_int32), ("bottom", C.c_int32)])
BorderScanParameters = _S("BorderScanParameters", [
    ("scan_size", RectangleSize), ("scan_step", Delta), ("scan_threshold", _I2),
    ("scan_direction", Direction)])

HostImage = _S("B200HostImage", [
    ("data", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32),
    ("linesize", C.c_int32), ("format", C.c_int32), ("background", Pixel),
    ("abs_black_threshold", C.c_uint8)])

MultiIndex = _S("B200MultiIndex", [("count", C.c_int32), ("indexes", C.POINTER(C.c_int32))])

SheetConfig = _S("B200SheetConfig", [
    ("layout", C.c_int32), ("input_count", C.c_int32), ("interpolate_type", C.c_int32),
    ("sheet_background", Pixel), ("mask_color", Pixel),
    ("abs_black_threshold", C.c_uint8), ("abs_white_threshold", C.c_uint8),
    ("no_blackfilter", C.c_uint8), ("no_noisefilter", C.c_uint8), ("no_blurfilter", C.c_uint8),
    ("no_grayfilter", C.c_uint8), ("no_mask_scan", C.c_uint8), ("no_mask_center", C.c_uint8),
    ("no_deskew", C.c_uint8), ("no_wipe", C.c_uint8), ("no_border", C.c_uint8),
    ("no_border_scan", C.c_uint8), ("no_border_align", C.c_uint8), ("reserved0", C.c_uint8),
    ("noisefilter_intensity", C.c_uint64),
    ("blackfilter", BlackfilterParameters), ("blurfilter", BlurfilterParameters),
    ("grayfilter", GrayfilterParameters), ("deskew", DeskewParameters),
    ("mask_detection", MaskDetectionParameters), ("mask_alignment", MaskAlignmentParameters),
    ("border_scan", BorderScanParameters),
    ("pre_border", Border), ("border", Border), ("post_border", Border),
    ("middle_wipe", C.c_int32 * 2),
    ("point_count", C.c_int32), ("points", Point * 8),
    ("pre_mask_count", C.c_int32), ("pre_masks", Rectangle * 8),
    ("pre_wipe_count", C.c_int32), ("wipe_count", C.c_int32), ("post_wipe_count", C.c_int32),
    ("pre_wipes", Rectangle * 8), ("wipes", Rectangle * 8), ("post_wipes", Rectangle * 8),
    ("pre_mirror", Direction), ("post_mirror", Direction), ("pre_shift", Delta), ("post_shift", Delta),
    ("output_count", C.c_int32), ("first_sheet_nr", C.c_int32),
    ("no_blackfilter_sheets", MultiIndex), ("no_noisefilter_sheets", MultiIndex), ("no_blurfilter_sheets", MultiIndex),
    ("no_grayfilter_sheets", MultiIndex), ("no_mask_scan_sheets", MultiIndex), ("no_mask_center_sheets", MultiIndex),
    ("no_deskew_sheets", MultiIndex), ("no_wipe_sheets", MultiIndex), ("no_border_sheets", MultiIndex),
    ("no_border_scan_sheets", MultiIndex), ("no_border_align_sheets", MultiIndex), ("ignore_sheets", MultiIndex),
    ("pre_rotate", C.c_int32), ("post_rotate", C.c_int32),
    ("sheet_size", RectangleSize), ("stretch_size", RectangleSize), ("page_size", RectangleSize),
    ("post_stretch_size", RectangleSize), ("post_page_size", RectangleSize),
    ("pre_zoom_factor", C.c_float), ("post_zoom_factor", C.c_float)])


def multi_index(values):
    """B200MultiIndex over a list of sheet numbers (None = all sheets); keeps the array alive."""
    m = MultiIndex()
    if values is None:
        m.count = -1
        return m
    arr = (C.c_int32 * max(len(values), 1))(*values)
    m.count, m.indexes = len(values), C.cast(arr, C.POINTER(C.c_int32))
    m._keep = arr
    return m

SheetResult = _S("B200SheetResult", [
    ("status", C.c_int32), ("sheet_width", C.c_int32), ("sheet_height", C.c_int32),
    ("deskew_mask_count", C.c_int32), ("deskew_masks", Rectangle * TRACE_MAX_MASKS),
    ("rotation", C.c_float * TRACE_MAX_MASKS),
    ("center_mask_count", C.c_int32), ("center_masks", Rectangle * TRACE_MAX_MASKS),
    ("centered", C.c_int32 * TRACE_MAX_MASKS),
    ("border_count", C.c_int32), ("borders", Border * MAX_PAGES),
    ("border_masks", Rectangle * MAX_PAGES),
    ("blackfilter_fills", C.c_int32), ("noise_clusters", C.c_int32),
    ("reserved", C.c_int32 * 6)])


def rect(x0, y0, x1, y1):
    r = Rectangle()
    r.vertex[0].x, r.vertex[0].y, r.vertex[1].x, r.vertex[1].y = x0, y0, x1, y1
    return r


def rect_tuple(r):
    return (r.vertex[0].x, r.vertex[0].y, r.vertex[1].x, r.vertex[1].y)


def border_tuple(b):
    return (b.left, b.top, b.right, b.bottom)


def bytes_per_row(fmt, width):
    return {FMT_GRAY8: width, FMT_Y400A: 2 * width, FMT_RGB24: 3 * width,
            FMT_MONOWHITE: (width + 7) // 8, FMT_MONOBLACK: (width + 7) // 8}[fmt]


def default_sheet_config():
    """Python restatement of unpaper_b200_sheet_config_defaults() (reference
    lib/options.c:22-170, src/cli/cli_options.c:229-274,:1108-1109)."""
    c = SheetConfig()
    c.layout, c.input_count, c.interpolate_type = LAYOUT_SINGLE, 1, INTERP_CUBIC
    c.sheet_background = Pixel(255, 255, 255)
    c.mask_color = Pixel(255, 255, 255)
    c.abs_black_threshold = int(255 * (1.0 - np.float32(0.33)))
    c.abs_white_threshold = int(255 * np.float32(0.9))
    c.noisefilter_intensity = 4
    bf = c.blackfilter
    bf.scan_size = RectangleSize(20, 20); bf.scan_step = Delta(5, 5)
    bf.scan_depth = _U2(500, 500); bf.scan_direction = Direction(True, True)
    bf.abs_threshold = int(np.float32(255) * np.float32(0.95)); bf.intensity = 20
    c.blurfilter = BlurfilterParameters(RectangleSize(100, 100), Delta(50, 50), 0.01)
    c.grayfilter = GrayfilterParameters(RectangleSize(50, 50), Delta(20, 20), int(255 * 0.5))
    # deskew.c:23: `float d` promoted to double, multiplied, rounded to float once
    d2r = lambda d: float(np.float32(float(np.float32(d)) * np.pi / 180.0))
    c.deskew = DeskewParameters(d2r(5.0), d2r(0.1), d2r(1.0), 1500, 0.5, Edges(True, False, True, False))
    c.mask_detection = MaskDetectionParameters(
        RectangleSize(50, 50), Delta(5, 5), _I2(-1, -1), Direction(True, False),
        _F2(0.1, 0.1), 100, -1, 100, -1)
    c.mask_alignment = MaskAlignmentParameters(Edges(False, False, False, False), Delta(0, 0))
    c.border_scan = BorderScanParameters(RectangleSize(5, 5), Delta(5, 5), _I2(5, 5), Direction(False, True))
    c.output_count, c.first_sheet_nr = 1, 1
    for f in ("sheet_size", "stretch_size", "page_size", "post_stretch_size", "post_page_size"):
        setattr(c, f, RectangleSize(-1, -1))
    c.pre_zoom_factor = c.post_zoom_factor = 1.0
    return c


_HOST_SIGS = {
    "wipe_rectangle": [C.POINTER(HostImage), C.POINTER(Rectangle), Pixel],
    "copy_rectangle": [C.POINTER(HostImage), C.POINTER(HostImage), C.POINTER(Rectangle), Point],
    "center_image": [C.POINTER(HostImage), C.POINTER(HostImage), Point, RectangleSize],
    "stretch": [C.POINTER(HostImage), C.POINTER(HostImage), C.c_int32],
    "resize": [C.POINTER(HostImage), C.POINTER(HostImage), C.c_int32],
    "flip_rotate_90": [C.POINTER(HostImage), C.POINTER(HostImage), C.c_int32],
    "mirror": [C.POINTER(HostImage), Direction],
    "shift": [C.POINTER(HostImage), C.POINTER(HostImage), Delta],
    "apply_masks": [C.POINTER(HostImage), C.POINTER(Rectangle), C.c_size_t, Pixel],
    "apply_wipes": [C.POINTER(HostImage), C.POINTER(Wipes), Pixel],
    "apply_border": [C.POINTER(HostImage), C.POINTER(Border), Pixel],
    "detect_masks": [C.POINTER(HostImage), C.POINTER(MaskDetectionParameters), C.POINTER(Point),
                     C.c_size_t, C.POINTER(Rectangle)],
    "center_mask": [C.POINTER(HostImage), Point, C.POINTER(Rectangle)],
    "align_mask": [C.POINTER(HostImage), C.POINTER(Rectangle), C.POINTER(Rectangle),
                   C.POINTER(MaskAlignmentParameters)],
    "detect_border": [C.POINTER(HostImage), C.POINTER(BorderScanParameters), C.POINTER(Rectangle),
                      C.POINTER(Border)],
    "blackfilter": [C.POINTER(HostImage), C.POINTER(BlackfilterParameters)],
    "blurfilter": [C.POINTER(HostImage), C.POINTER(BlurfilterParameters), C.c_uint8],
    "noisefilter": [C.POINTER(HostImage), C.c_uint64, C.c_uint8],
    "grayfilter": [C.POINTER(HostImage), C.POINTER(GrayfilterParameters)],
    "detect_rotation": [C.POINTER(HostImage), C.POINTER(Rectangle), C.POINTER(DeskewParameters),
                        C.POINTER(C.c_float)],
    "deskew": [C.POINTER(HostImage), C.POINTER(Rectangle), C.c_float, C.c_int32],
}


_OUTPUT_SIGS = {
    "convert_format": [C.POINTER(HostImage), C.POINTER(HostImage)],
}


class HostOps:
    """The 21 host-buffer ops of one library, numpy in / numpy out.

    Images are ``np.uint8`` arrays of shape (H, linesize); ``fmt`` says how a
    row is to be read.  In-place ops modify the array that is passed in.
    """

    def __init__(self, lib, prefix):
        self.lib, self.prefix = lib, prefix
        self.fn = {}
        for name, sig in _HOST_SIGS.items():
            f = getattr(lib, prefix + name)
            f.argtypes, f.restype = sig, C.c_int
            self.fn[name] = f
        # output side: exists in the product and in the restatement, not in the
        # reference build (file.c needs libavcodec)
        for name, sig in _OUTPUT_SIGS.items():
            if hasattr(lib, prefix + name):
                f = getattr(lib, prefix + name)
                f.argtypes, f.restype = sig, C.c_int
                self.fn[name] = f

    @staticmethod
    def himg(a, fmt, width, bg=(255, 255, 255), abt=170):
        assert a.dtype == np.uint8 and a.ndim == 2 and a.flags["C_CONTIGUOUS"]
        assert a.shape[1] >= bytes_per_row(fmt, width)
        return HostImage(a.ctypes.data, width, a.shape[0], a.shape[1], fmt, Pixel(*bg), abt)

    def _chk(self, rc, name):
        if rc < 0:
            raise RuntimeError(f"{self.prefix}{name} failed: {rc}")
        return rc

    def call(self, name, *args):
        return self._chk(self.fn[name](*args), name)
