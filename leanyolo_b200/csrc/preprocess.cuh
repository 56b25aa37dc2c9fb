// cv2.resize(INTER_LINEAR) of 8-bit images + constant border, restated bit for bit (OpenCV imgproc/resize.cpp,
// 11-bit fixed-point coefficients; oracle/letterbox_oracle.py has the arithmetic).  Shared by the stand-alone
// letterbox kernel (preprocess.cu) and the stem kernel's fused loader (elementwise.cu).
#pragma once
#include "common.cuh"

namespace ly {

// one axis of cv::resize's coefficient table
struct AxisCoef { int s; int c0, c1; };

__device__ __forceinline__ AxisCoef axis_coef(int d, int dst, int src, bool clamp_fraction) {
  const double scale = __drcp_rn(__ddiv_rn((double)dst, (double)src));       // 1 / (dst / src)
  const float f0 = (float)__dadd_rn(__dmul_rn((double)d + 0.5, scale), -0.5);
  int s = (int)floorf(f0);
  float f = __fsub_rn(f0, (float)s);
  if (clamp_fraction) {
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= src - 1) { s = src - 1; f = 0.f; }
  }
  AxisCoef a;
  a.s = s;
  a.c0 = __float2int_rn(__fmul_rn(__fsub_rn(1.0f, f), 2048.0f));
  a.c1 = __float2int_rn(__fmul_rn(f, 2048.0f));
  return a;
}


// resize class of a letterbox descriptor: 0 plain copy, 1 exact 2x decimation (cv::resize switches to INTER_AREA), 2 bilinear
__device__ __forceinline__ int lb_mode(const ly_lb_desc& d) {
  if (d.new_w == d.src_w && d.new_h == d.src_h) return 0;
  if (d.src_w == 2 * d.new_w && d.src_h == 2 * d.new_h) return 1;
  return 2;
}

// bilinear sample with precomputed axis coefficients (x axis: fraction clamped at the borders; y axis: indices clipped)
__device__ __forceinline__ void lb_bilinear(const ly_lb_desc& d, const AxisCoef& ax, const AxisCoef& ay, int* v) {
  const int x0 = ax.s, x1 = min(ax.s + 1, d.src_w - 1);
  const int y0 = min(max(ay.s, 0), d.src_h - 1), y1 = min(max(ay.s + 1, 0), d.src_h - 1);
  const uint8_t* r0 = d.src + y0 * d.src_pitch;
  const uint8_t* r1 = d.src + y1 * d.src_pitch;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int S0 = r0[3 * x0 + c] * ax.c0 + r0[3 * x1 + c] * ax.c1;
    const int S1 = r1[3 * x0 + c] * ax.c0 + r1[3 * x1 + c] * ax.c1;
    v[c] = (((ay.c0 * (S0 >> 4)) >> 16) + ((ay.c1 * (S1 >> 4)) >> 16) + 2) >> 2;
  }
}

// pixel (x, y) of the letterboxed image: border colour outside the resized picture
__device__ __forceinline__ void lb_sample(const ly_lb_desc& d, int x, int y, int fr, int fg, int fb, int* v) {
  v[0] = fr; v[1] = fg; v[2] = fb;
  const int rx = x - d.left, ry = y - d.top;
  if (rx < 0 || rx >= d.new_w || ry < 0 || ry >= d.new_h) return;
  const int mode = lb_mode(d);
  if (mode == 0) {
    const uint8_t* p = d.src + ry * d.src_pitch + 3 * rx;
    v[0] = p[0]; v[1] = p[1]; v[2] = p[2];
  } else if (mode == 1) {
    const uint8_t* p0 = d.src + (2 * ry) * d.src_pitch + 6 * rx;
    const uint8_t* p1 = p0 + d.src_pitch;
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = (p0[c] + p0[3 + c] + p1[c] + p1[3 + c] + 2) >> 2;
  } else {
    lb_bilinear(d, axis_coef(rx, d.new_w, d.src_w, true), axis_coef(ry, d.new_h, d.src_h, false), v);
  }
}

}  // namespace ly
