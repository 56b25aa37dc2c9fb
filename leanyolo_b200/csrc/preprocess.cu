// Pre/post-processing on either side of the hot path.
//
//  * letterbox_kernel: uint8 HWC image -> aspect-preserving bilinear resize + constant border ->
//    one [3, S, S] (CHW) or [S, S, 3] (HWC) uint8 slot of the batch the stem kernel consumes
//    (reference: leanyolo/utils/letterbox.py:9-91 + the HWC->CHW transpose of tools/infer.py:112-114).
//    The resize is cv2.resize(..., INTER_LINEAR) of 8-bit images restated bit for bit (OpenCV
//    imgproc/resize.cpp, 11-bit fixed-point coefficients; see oracle/letterbox_oracle.py for the
//    arithmetic): every operation below is an IEEE op in the same order (no FMA contraction in the
//    coordinate computation), so the GPU result equals cv2's.
//  * unletterbox_kernel: detections back to original-image coordinates, in place
//    (leanyolo/utils/box_ops.py:96-124): (x - pad) / gain, clamped to the image.
//
// Both are pure bandwidth work: one thread per output pixel (3 channels) / per box.
#include "common.cuh"
#include "preprocess.cuh"

namespace ly {

namespace {

__global__ void __launch_bounds__(256) letterbox_kernel(const ly_lb_desc* __restrict__ descs, uint8_t* __restrict__ dst, int dst_h,
                                                        int dst_w, int chw, int fr, int fg, int fb) {
  const ly_lb_desc d = descs[blockIdx.z];
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= dst_w || y >= dst_h) return;
  int v[3];
  lb_sample(d, x, y, fr, fg, fb, v);
  uint8_t* o = dst + (size_t)blockIdx.z * 3 * dst_h * dst_w;
  if (chw) {
    const size_t plane = (size_t)dst_h * dst_w, at = (size_t)y * dst_w + x;
    o[at] = (uint8_t)v[0]; o[plane + at] = (uint8_t)v[1]; o[2 * plane + at] = (uint8_t)v[2];
  } else {
    uint8_t* q = o + ((size_t)y * dst_w + x) * 3;
    q[0] = (uint8_t)v[0]; q[1] = (uint8_t)v[1]; q[2] = (uint8_t)v[2];
  }
}

__global__ void __launch_bounds__(256) unletterbox_kernel(float* __restrict__ dets, int B, int K, int row, const float* __restrict__ meta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * K) return;
  const int b = i / K;
  const float gw = meta[6 * b + 0], gh = meta[6 * b + 1], px = meta[6 * b + 2], py = meta[6 * b + 3];
  const float H = meta[6 * b + 4], W = meta[6 * b + 5];
  float* d = dets + (size_t)i * row;
  d[0] = fminf(fmaxf(__fdiv_rn(__fsub_rn(d[0], px), gw), 0.f), W);
  d[1] = fminf(fmaxf(__fdiv_rn(__fsub_rn(d[1], py), gh), 0.f), H);
  d[2] = fminf(fmaxf(__fdiv_rn(__fsub_rn(d[2], px), gw), 0.f), W);
  d[3] = fminf(fmaxf(__fdiv_rn(__fsub_rn(d[3], py), gh), 0.f), H);
}

}  // namespace

int32_t launch_letterbox(const ly_lb_desc* descs, int32_t B, uint8_t* dst, int32_t dst_h, int32_t dst_w, int32_t chw,
                         const uint8_t* fill, cudaStream_t s) {
  LY_CHECK_ARG(descs && dst && B > 0 && dst_h > 0 && dst_w > 0 && B <= 65535 && dst_h <= 65535, "letterbox: bad arguments");
  dim3 grid((dst_w + 255) / 256, dst_h, B);
  letterbox_kernel<<<grid, 256, 0, s>>>(descs, dst, dst_h, dst_w, chw, fill ? fill[0] : 114, fill ? fill[1] : 114, fill ? fill[2] : 114);
  return post_launch("letterbox");
}

int32_t launch_unletterbox(float* dets, int32_t B, int32_t K, int32_t row, const float* meta, cudaStream_t s) {
  LY_CHECK_ARG(dets && meta && B > 0 && K > 0 && row >= 4, "unletterbox: bad arguments");
  unletterbox_kernel<<<(B * K + 255) / 256, 256, 0, s>>>(dets, B, K, row, meta);
  return post_launch("unletterbox");
}

}  // namespace ly
