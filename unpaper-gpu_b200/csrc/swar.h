/* swar.h — byte compares on four packed pixels without the SIMD-video intrinsics.
 *
 * bit 7 of every byte of lt4(w, lt4_make(T)) is set where that byte of w is < T, every other bit is clear.
 *   (w & 0x7F7F7F7F) + (0x80 - (T & 0x7F)) per byte never carries into the next byte and has bit 7 set
 *   iff the low seven bits of the byte are >= those of T; with the top bit of the byte itself:
 *   byte >= T  <=>  T < 128 ? (top | low_ge) : (top & low_ge).
 * Three integer instructions per bound (LOP3, IADD, LOP3) instead of the six __vcmpltu4 / __vcmpgeu4 expand to.
 * Compiles as C as well, so that the CPU suite can check it over every (threshold, byte) pair
 * (tests/test_abi.py::test_swar_byte_compare). */
#pragma once
#ifdef __CUDACC__
#define SWAR_FN __host__ __device__ __forceinline__
#else
#define SWAR_FN static inline
#endif

typedef struct { unsigned addv, both; } Lt4;   /* both: all ones for T >= 128 (top & low_ge), zero for T < 128 (top | low_ge) */

SWAR_FN Lt4 lt4_make(int T) {
  Lt4 r;
  T = T < 0 ? 0 : T;                        /* nothing is below 0: the sum always has bit 7 set, every byte is ">= T" */
  r.both = T >= 128 ? 0xFFFFFFFFu : 0u;
  r.addv = T >= 256 ? 0u : (0x80u - ((unsigned)T & 0x7Fu)) * 0x01010101u;   /* T >= 256: bit 7 of the sum never set, no byte is ">= T" */
  return r;
}
SWAR_FN unsigned lt4(unsigned w, Lt4 t) {
  const unsigned s = (w & 0x7F7F7F7Fu) + t.addv;
  const unsigned ge = (w & s) | ((w | s) & ~t.both);
  return ~ge & 0x80808080u;
}
/* bit 7 of every byte set where lo <= byte <= hi (hiT = lt4_make(hi + 1), loT = lt4_make(lo)) */
SWAR_FN unsigned range4_bit7(unsigned w, Lt4 loT, Lt4 hiT) { return lt4(w, hiT) & ~lt4(w, loT); }
