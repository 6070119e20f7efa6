// conv_simt.cu -- generic fp32 CUDA-core convolution on NHWC activations with OIHW weights.
// Covers what the tensor-core kernels do not take: the 7x7 stride-2 stem (reference
// lib/models/hourglass.py:116, Cin=3), the 16-channel score / score_ 1x1 convs (hourglass.py:135-137)
// and odd shapes; it is also the on-device cross-check the tests use for conv_tc / wgrad_tc.
// fp32 FFMA accumulation throughout, so results track the fp32 reference to round-off.
#include "common.cuh"
#include "kernels.h"

namespace fpd {
namespace {

// Each thread: one output pixel x 4 consecutive output channels. Threads of a warp share the pixel
// group's inputs through L1; weights are read from the OIHW tensor through the read-only path.
__global__ void __launch_bounds__(256)
conv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                const float* __restrict__ residual, float* __restrict__ y, int B, int H, int W, int Cin, int Cout,
                int k, int stride, int pad, int Ho, int Wo) {
  const int cq = (Cout + 3) / 4;
  const int64_t n = (int64_t)B * Ho * Wo * cq;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % cq) * 4;
    int64_t t = i / cq;
    const int wo = (int)(t % Wo); t /= Wo;
    const int ho = (int)(t % Ho);
    const int b = (int)(t / Ho);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const int taps = k * k;
    for (int kh = 0; kh < k; ++kh) {
      const int hi = ho * stride + kh - pad;
      if (hi < 0 || hi >= H) continue;
      for (int kw = 0; kw < k; ++kw) {
        const int wi = wo * stride + kw - pad;
        if (wi < 0 || wi >= W) continue;
        const float* xp = x + (((int64_t)b * H + hi) * W + wi) * Cin;
        const int tap = kh * k + kw;
        for (int ci = 0; ci < Cin; ++ci) {
          const float xv = __ldg(xp + ci);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (c4 + j < Cout) acc[j] = fmaf(xv, __ldg(w + ((int64_t)(c4 + j) * Cin + ci) * taps + tap), acc[j]);
          }
        }
      }
    }
    const int64_t o = (((int64_t)b * Ho + ho) * Wo + wo) * Cout + c4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (c4 + j < Cout) {
        float v = acc[j];
        if (bias) v += __ldg(bias + c4 + j);
        if (residual) v += __ldg(residual + o + j);
        y[o + j] = v;
      }
    }
  }
}

// dx[b,hi,wi,ci] = sum_{co,kh,kw : ho*stride+kh-pad==hi ...} dy[b,ho,wo,co] * w[co,ci,kh,kw]
__global__ void __launch_bounds__(256)
conv_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx, int B, int H,
                  int W, int Cin, int Cout, int k, int stride, int pad, int Ho, int Wo) {
  const int cq = (Cin + 3) / 4;
  const int64_t n = (int64_t)B * H * W * cq;
  const int taps = k * k;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % cq) * 4;
    int64_t t = i / cq;
    const int wi = (int)(t % W); t /= W;
    const int hi = (int)(t % H);
    const int b = (int)(t / H);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int kh = 0; kh < k; ++kh) {
      const int hn = hi + pad - kh;
      if (hn < 0 || hn % stride != 0) continue;
      const int ho = hn / stride;
      if (ho >= Ho) continue;
      for (int kw = 0; kw < k; ++kw) {
        const int wn = wi + pad - kw;
        if (wn < 0 || wn % stride != 0) continue;
        const int wo = wn / stride;
        if (wo >= Wo) continue;
        const float* gp = dy + (((int64_t)b * Ho + ho) * Wo + wo) * Cout;
        const int tap = kh * k + kw;
        for (int co = 0; co < Cout; ++co) {
          const float g = __ldg(gp + co);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (c4 + j < Cin) acc[j] = fmaf(g, __ldg(w + ((int64_t)co * Cin + c4 + j) * taps + tap), acc[j]);
          }
        }
      }
    }
    const int64_t o = (((int64_t)b * H + hi) * W + wi) * Cin + c4;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (c4 + j < Cin) dx[o + j] = acc[j];
  }
}

// dw[co,ci,kh,kw] = scale * sum_{b,ho,wo} dy[b,ho,wo,co] * x[b,ho*s+kh-p,wo*s+kw-p,ci]
// grid: (taps*Cin, ceil(Cout/32)); block 256 threads stride over output pixels; each thread keeps 32
// output-channel partials in registers; fixed-order block reduction (deterministic).
__global__ void __launch_bounds__(256)
conv_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw, float scale,
                  int B, int H, int W, int Cin, int Cout, int k, int stride, int pad, int Ho, int Wo) {
  const int taps = k * k;
  const int tap = blockIdx.x / Cin, ci = blockIdx.x % Cin;
  const int kh = tap / k, kw = tap % k;
  const int co0 = blockIdx.y * 32;
  const int nco = min(32, Cout - co0);
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.f;
  const int64_t P = (int64_t)B * Ho * Wo;
  for (int64_t p = threadIdx.x; p < P; p += blockDim.x) {
    const int wo = (int)(p % Wo);
    const int ho = (int)((p / Wo) % Ho);
    const int b = (int)(p / ((int64_t)Wo * Ho));
    const int hi = ho * stride + kh - pad, wi = wo * stride + kw - pad;
    if (hi < 0 || hi >= H || wi < 0 || wi >= W) continue;
    const float xv = __ldg(x + (((int64_t)b * H + hi) * W + wi) * Cin + ci);
    const float* gp = dy + p * Cout + co0;
    if (nco == 32 && (Cout % 4 == 0)) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gp + j));
        acc[j] = fmaf(xv, g.x, acc[j]);
        acc[j + 1] = fmaf(xv, g.y, acc[j + 1]);
        acc[j + 2] = fmaf(xv, g.z, acc[j + 2]);
        acc[j + 3] = fmaf(xv, g.w, acc[j + 3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nco) acc[j] = fmaf(xv, __ldg(gp + j), acc[j]);
    }
  }
  __shared__ float red[8][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float s = warp_sum(acc[j]);
    if (lane == 0) red[wid][j] = s;
  }
  __syncthreads();
  if (threadIdx.x < nco) {
    float s = 0.f;
    for (int wv = 0; wv < 8; ++wv) s += red[wv][threadIdx.x];
    dw[((int64_t)(co0 + threadIdx.x) * Cin + ci) * taps + tap] = s * scale;
  }
}

inline int grid_for(int64_t n, int threads) {
  int64_t b = (n + threads - 1) / threads;
  const int64_t cap = 148 * 32;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace

int conv_simt_fwd(const float* x, const float* w_oihw, const float* bias, const float* residual, float* y, int B,
                  int H, int W, int Cin, int Cout, int k, int stride, int pad, cudaStream_t stream) {
  FPD_REQUIRE(k >= 1 && stride >= 1 && pad >= 0, "conv_simt_fwd: bad geometry");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const int64_t n = (int64_t)B * Ho * Wo * ((Cout + 3) / 4);
  conv_fwd_kernel<<<grid_for(n, 256), 256, 0, stream>>>(x, w_oihw, bias, residual, y, B, H, W, Cin, Cout, k, stride,
                                                        pad, Ho, Wo);
  FPD_LAUNCH_CHECK();
  return FPD_OK;
}

int conv_simt_dgrad(const float* dy, const float* w_oihw, float* dx, int B, int H, int W, int Cin, int Cout, int k,
                    int stride, int pad, cudaStream_t stream) {
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const int64_t n = (int64_t)B * H * W * ((Cin + 3) / 4);
  conv_dgrad_kernel<<<grid_for(n, 256), 256, 0, stream>>>(dy, w_oihw, dx, B, H, W, Cin, Cout, k, stride, pad, Ho,
                                                          Wo);
  FPD_LAUNCH_CHECK();
  return FPD_OK;
}

int conv_simt_wgrad(const float* x, const float* dy, float* dw_oihw, float scale, int B, int H, int W, int Cin,
                    int Cout, int k, int stride, int pad, cudaStream_t stream) {
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  dim3 grid(k * k * Cin, (Cout + 31) / 32);
  conv_wgrad_kernel<<<grid, 256, 0, stream>>>(x, dy, dw_oihw, scale, B, H, W, Cin, Cout, k, stride, pad, Ho, Wo);
  FPD_LAUNCH_CHECK();
  return FPD_OK;
}

}  // namespace fpd
