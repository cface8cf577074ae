// Crop -> bicubic resize -> normalise, on the device (SURVEY.md §8 b1 + §8f N3).
//
// Replaces, for objects cut out of frames that are already in HBM for the detector: the crop of
// `VideoCapture.save_object` (clearcam.py:396), the BGR->RGB swap (models/objects.py:249) and `ObjectFinder.preprocess`
// (models/objects.py:237-242: cv2.resize(INTER_CUBIC) -> /255 -> (x-0.5)/0.5 -> CHW), skipping the reference's JPEG
// write/read in between.  Bit-exact with OpenCV's own 8-bit bicubic (see oracle/clip_preprocess.py for the arithmetic
// and its pinning against cv2): taps in float32 with one rounding per operation (hence the explicit _rn intrinsics —
// no fused multiply-add), 11-bit fixed-point taps, exact int32 horizontal pass, float32 vertical pass, round-half-even.
//
// HBM-bound and tiny next to the encoder (K x 224 x 224 outputs, 48 byte reads each, served from L1/L2): one thread per
// output pixel, three channels, coalesced plane writes.
#include <cstdint>

#include "cc_common.h"
#include "clearcam_b200.h"
#include "ops.cuh"

namespace cc {
namespace {

constexpr int kRectsPerLaunch = 64;
struct CropParams {
  const uint8_t* frames;
  float* out;
  int H, W, S, bgr, k0;
  int rect[kRectsPerLaunch][5];     // frame, x1, y1, x2, y2 (validated on the host)
};

// source index of the first tap (+1) and the four fixed-point taps for destination index d (resize.cpp, a = -0.75)
__device__ __forceinline__ void cubic_taps(int d, int src, int dst, int& s, int taps[4]) {
  const double scale = __ddiv_rn(1.0, __ddiv_rn(static_cast<double>(dst), static_cast<double>(src)));
  const float f = static_cast<float>(__dsub_rn(__dmul_rn(__dadd_rn(static_cast<double>(d), 0.5), scale), 0.5));
  const float fl = floorf(f);
  s = static_cast<int>(fl);
  const float x = __fsub_rn(f, fl);
  const float x1 = __fadd_rn(x, 1.f), xm = __fsub_rn(1.f, x);
  // 5A = -3.75, 8A = -6, 4A = -3, A+2 = 1.25, A+3 = 2.25
  const float c0 = __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(__fmul_rn(-0.75f, x1), -3.75f), x1), -6.f), x1), -3.f);
  const float c1 = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(1.25f, x), 2.25f), x), x), 1.f);
  const float c2 = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(1.25f, xm), 2.25f), xm), xm), 1.f);
  const float c3 = __fsub_rn(__fsub_rn(__fsub_rn(1.f, c0), c1), c2);
  taps[0] = __float2int_rn(__fmul_rn(c0, 2048.f));
  taps[1] = __float2int_rn(__fmul_rn(c1, 2048.f));
  taps[2] = __float2int_rn(__fmul_rn(c2, 2048.f));
  taps[3] = __float2int_rn(__fmul_rn(c3, 2048.f));
}

__global__ void __launch_bounds__(256) crop_resize_kernel(const __grid_constant__ CropParams p) {
  const int dx = blockIdx.x * 32 + threadIdx.x, dy = blockIdx.y * 8 + threadIdx.y, k = blockIdx.z;
  if (dx >= p.S || dy >= p.S) return;
  const int fr = p.rect[k][0], x1 = p.rect[k][1], y1 = p.rect[k][2];
  const int cw = p.rect[k][3] - x1, ch = p.rect[k][4] - y1;
  const uint8_t* img = p.frames + (static_cast<size_t>(fr) * p.H + y1) * p.W * 3 + static_cast<size_t>(x1) * 3;
  int sx, sy, ax[4], ay[4];
  cubic_taps(dx, cw, p.S, sx, ax);
  cubic_taps(dy, ch, p.S, sy, ay);
  int hor[4][3];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int yy = min(max(sy - 1 + r, 0), ch - 1);
    const uint8_t* row = img + static_cast<size_t>(yy) * p.W * 3;
    int acc[3] = {0, 0, 0};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int xx = min(max(sx - 1 + j, 0), cw - 1);
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[c] += static_cast<int>(__ldg(row + xx * 3 + c)) * ax[j];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) hor[r][c] = acc[c];
  }
  const float kInv = 1.0f / (2048.f * 2048.f);
  float b[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) b[r] = __fmul_rn(static_cast<float>(ay[r]), kInv);
  const int tail = (p.S * 3) / 8 * 8;            // interleaved elements past the last group of 8 take the integer form
  float* out = p.out + static_cast<size_t>(p.k0 + k) * 3 * p.S * p.S + static_cast<size_t>(dy) * p.S + dx;
#pragma unroll
  for (int c = 0; c < 3; ++c) {                  // c indexes the crop's channel order after the optional BGR->RGB swap
    const int sc = p.bgr ? 2 - c : c;
    int v;
    if (dx * 3 + c < tail) {
      float t = __fmul_rn(static_cast<float>(hor[3][sc]), b[3]);
      t = __fadd_rn(__fmul_rn(static_cast<float>(hor[2][sc]), b[2]), t);
      t = __fadd_rn(__fmul_rn(static_cast<float>(hor[1][sc]), b[1]), t);
      t = __fadd_rn(__fmul_rn(static_cast<float>(hor[0][sc]), b[0]), t);
      v = __float2int_rn(t);
    } else {
      const long long s = static_cast<long long>(hor[0][sc]) * ay[0] + static_cast<long long>(hor[1][sc]) * ay[1] +
                          static_cast<long long>(hor[2][sc]) * ay[2] + static_cast<long long>(hor[3][sc]) * ay[3];
      v = static_cast<int>((s + (1 << 21)) >> 22);
    }
    v = min(max(v, 0), 255);
    const float n = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(v), 255.f), 0.5f), 0.5f);     // models/objects.py:239-240
    out[static_cast<size_t>(c) * p.S * p.S] = n;
  }
}

}  // namespace
}  // namespace cc

extern "C" int cc_clip_preprocess(const uint8_t* d_frames, int n_frames, int H, int W, const int32_t* rects, int K, int size,
                                  int bgr, float* d_out, void* stream) {
  using namespace cc;
  CC_REQUIRE(d_frames && d_out && (rects || K == 0), "cc_clip_preprocess: null pointer");
  CC_REQUIRE(n_frames > 0 && H > 0 && W > 0 && K >= 0 && size >= 1 && size <= 4096, "cc_clip_preprocess: bad sizes");
  for (int k = 0; k < K; ++k) {
    const int32_t* r = rects + 5 * k;
    CC_REQUIRE(r[0] >= 0 && r[0] < n_frames && r[1] >= 0 && r[2] >= 0 && r[3] > r[1] && r[4] > r[2] && r[3] <= W && r[4] <= H,
               "cc_clip_preprocess: rect %d (frame %d, %d,%d-%d,%d) is empty or outside the %dx%d frame", k, r[0], r[1], r[2],
               r[3], r[4], W, H);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int k0 = 0; k0 < K; k0 += kRectsPerLaunch) {
    CropParams p;
    p.frames = d_frames;
    p.out = d_out;
    p.H = H;
    p.W = W;
    p.S = size;
    p.bgr = bgr;
    p.k0 = k0;
    const int n = K - k0 < kRectsPerLaunch ? K - k0 : kRectsPerLaunch;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < 5; ++j) p.rect[i][j] = rects[5 * (k0 + i) + j];
    crop_resize_kernel<<<dim3((size + 31) / 32, (size + 7) / 8, n), dim3(32, 8), 0, st>>>(p);
    CC_CHECK_CUDA(cudaGetLastError());
  }
  return CC_OK;
}
