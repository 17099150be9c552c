// pg_math.h — the fp64 / fixed-point arithmetic shared by every kernel, written once as
// host+device inline functions so the CPU test-suite can evaluate the very same source
// (pg_hostcheck_* in pagegeom.h) against the oracle before any GPU time is spent.
//
// Everything here mirrors the reference's Python-double operation order; the library is
// compiled with -fmad=false so neither nvcc nor gcc contracts a*b+c into an FMA.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PG_HD __host__ __device__ __forceinline__
#else
#define PG_HD inline
#endif

// calculate_iou (3_combine_grids.py:46-78).  `area_a`/`area_b` are (x1-x0)*(y1-y0),
// precomputed once per box (same expression, same rounding).
PG_HD double pg_iou(double ax0, double ay0, double ax1, double ay1, double area_a,
                    double bx0, double by0, double bx1, double by1, double area_b) {
  const double xl = ax0 > bx0 ? ax0 : bx0;   // max(box1[0], box2[0])
  const double yt = ay0 > by0 ? ay0 : by0;
  const double xr = ax1 < bx1 ? ax1 : bx1;   // min(box1[2], box2[2])
  const double yb = ay1 < by1 ? ay1 : by1;
  if (xr < xl || yb < yt) return 0.0;        // :65 (touching boxes fall through to area 0)
  const double inter = (xr - xl) * (yb - yt);
  const double uni = area_a + area_b - inter;  // :73, left to right
  return uni > 0.0 ? inter / uni : 0.0;        // :76, IEEE divide
}

PG_HD double pg_box_area(double x0, double y0, double x1, double y1) { return (x1 - x0) * (y1 - y0); }

// Exactly `pg_iou(a, b) > thr` (3_combine_grids.py:130), without the divide on the common path.
// With p = fl(thr*uni) (relative error <= 2^-53): inter > fl(p*(1+2^-50)) puts the real quotient
// above thr*(1+2^-51), i.e. beyond the midpoint to nextafter(thr), so fl(inter/uni) > thr;
// inter < fl(p*(1-2^-50)) puts it below thr, so fl(inter/uni) <= thr.  Only inside that 2^-50-wide
// band is the IEEE divide evaluated (thr > 0 and normal-range operands; every other case falls
// back to the literal expression).
PG_HD bool pg_iou_gt(double ax0, double ay0, double ax1, double ay1, double area_a,
                     double bx0, double by0, double bx1, double by1, double area_b, double thr) {
  const double xl = ax0 > bx0 ? ax0 : bx0;
  const double yt = ay0 > by0 ? ay0 : by0;
  const double xr = ax1 < bx1 ? ax1 : bx1;
  const double yb = ay1 < by1 ? ay1 : by1;
  if (xr < xl || yb < yt) return 0.0 > thr;
  const double inter = (xr - xl) * (yb - yt);
  const double uni = area_a + area_b - inter;
  if (!(uni > 0.0)) return 0.0 > thr;
  if (thr > 1e-300 && thr < 1e300 && uni > 1e-280 && uni < 1e280) {
    const double p = thr * uni;
    if (inter > p * (1.0 + 0x1p-50)) return true;
    if (inter < p * (1.0 - 0x1p-50)) return false;
  }
  return inter / uni > thr;
}

// torchvision.ops.nms's test (called at 1_doclayout_bboxes.py:219-223 on float32 boxes): everything in
// float32 — w = max(0, xx2-xx1), h likewise, inter = w*h, ovr = inter / (area_i + area_j - inter) —
// then `ovr > iou_threshold` with the threshold a double.  0/0 gives NaN, which compares false.
PG_HD float pg_box_area_f32(float x0, float y0, float x1, float y1) { return (x1 - x0) * (y1 - y0); }
PG_HD bool pg_iou_gt_f32(float ix0, float iy0, float ix1, float iy1, float area_i,
                         float jx0, float jy0, float jx1, float jy1, float area_j, double thr) {
  const float xx1 = ix0 > jx0 ? ix0 : jx0;
  const float yy1 = iy0 > jy0 ? iy0 : jy0;
  const float xx2 = ix1 < jx1 ? ix1 : jx1;
  const float yy2 = iy1 < jy1 ? iy1 : jy1;
  const float dw = xx2 - xx1, dh = yy2 - yy1;
  const float w = dw > 0.f ? dw : 0.f;
  const float h = dh > 0.f ? dh : 0.f;
  const float inter = w * h;
  const float ovr = inter / (area_i + area_j - inter);
  return (double)ovr > thr;
}

// is_box_touching_internal_edge (2_edge_box_filter.py:44-90), test order right, bottom,
// left, top.  Pure predicate: the order only matters for short-circuiting.
PG_HD bool pg_edge_touch(double x_min, double y_min, double x_max, double y_max,
                         double cx0, double cy0, double cx1, double cy1,
                         double image_w, double image_h, double thr) {
  if (fabs(cx1 - image_w) > thr && x_max >= (cx1 - thr)) return true;   // :71-73
  if (fabs(cy1 - image_h) > thr && y_max >= (cy1 - thr)) return true;   // :76-78
  if (cx0 > thr && x_min <= (cx0 + thr)) return true;                   // :81-83
  if (cy0 > thr && y_min <= (cy0 + thr)) return true;                   // :86-88
  return false;
}

// Python floor division for a positive divisor.
PG_HD int32_t pg_floordiv(int32_t a, int32_t b) {
  int32_t q = a / b;
  return (a % b != 0 && a < 0) ? q - 1 : q;
}

// density weight of one bin inside one box span (5_detect_column_centers.py:140-143).
PG_HD double pg_density_weight(int32_t bin, int32_t left, int32_t right, int32_t center) {
  const int32_t ad = bin - center;
  const double dist = (double)(ad < 0 ? -ad : ad) / ((double)(right - left) / 2.0 + 1e-6);
  return 1.0 - 0.5 * (dist < 1.0 ? dist : 1.0);
}

// Same value with the divide replaced by a reciprocal and one FMA correction (Markstein):
//   y = RN(1/half) (one IEEE divide per box), q0 = n*y, r = fma(-q0, half, n), dist = fma(r, y, q0).
// dist is the correctly rounded n/half; verified exhaustively (host test) for every integer
// n, right-left in [0, PG_RCP_DOMAIN], which covers all density bins (< 2048).
#define PG_RCP_DOMAIN 2100
PG_HD double pg_density_half(int32_t left, int32_t right) { return (double)(right - left) / 2.0 + 1e-6; }
PG_HD double pg_density_weight_rcp(int32_t bin, int32_t center, double half, double inv_half) {
  const int32_t ad = bin - center;
  const double n = (double)(ad < 0 ? -ad : ad);
  const double q0 = n * inv_half;
  const double rem = fma(-q0, half, n);
  const double dist = fma(rem, inv_half, q0);
  return 1.0 - 0.5 * (dist < 1.0 ? dist : 1.0);
}

// ---------------------------------------------------------------------------------------------
// cv2.resize(uint8, INTER_LINEAR) fixed-point model (validated bit-exact against
// cv2 4.13).  Coefficient for destination index d: s0/s1 = the two source indices, c0/c1 =
// 11-bit weights.  x clamps the fraction at the borders, y clamps the indices only.
struct PgCoef {
  int32_t s0, s1;
  int32_t c0, c1;
};

inline PgCoef pg_resize_coef(int32_t ssize, int32_t dsize, int32_t d, bool clamp_frac) {
  const double inv = (double)dsize / (double)ssize;
  const double scale = 1.0 / inv;
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int32_t s = (int32_t)floorf(f);
  f -= (float)s;
  if (clamp_frac) {
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
  }
  PgCoef c;
  c.c0 = (int32_t)nearbyintf((1.f - f) * 2048.f);
  c.c1 = (int32_t)nearbyintf(f * 2048.f);
  c.s0 = s < 0 ? 0 : (s > ssize - 1 ? ssize - 1 : s);
  c.s1 = s + 1 < 0 ? 0 : (s + 1 > ssize - 1 ? ssize - 1 : s + 1);
  return c;
}

// horizontal pass for one channel: S0*a0 + S1*a1 (fits 20 bits)
PG_HD uint32_t pg_hpass(uint32_t s0, uint32_t s1, uint32_t a0, uint32_t a1) { return s0 * a0 + s1 * a1; }
// vertical pass: (((b0*(h0>>4))>>16) + ((b1*(h1>>4))>>16) + 2) >> 2
// (Measured and rejected: the equivalent mulhi32(b<<12, h & ~15) form — IMAD.HI is slow enough on
// sm_100a that the tiler went from 2.21 to 2.31 ms per 64 pages.)
PG_HD uint32_t pg_vpass(uint32_t h0, uint32_t h1, uint32_t b0, uint32_t b1) {
  return (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2u) >> 2;
}
