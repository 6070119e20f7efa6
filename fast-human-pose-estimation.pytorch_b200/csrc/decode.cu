// decode.cu -- device-side flip-test merge and heat-map arg-max.
//
// Reference (host numpy, three host round-trips per batch):
//   lib/core/function.py:218-240  flip the input, second forward, flip_back, 1-px shift, average
//   lib/utils/transforms.py:15-29 flip_back: reverse W, swap left/right joint channels
//   lib/core/inference.py:18-46   get_max_preds: flat arg-max per (b,j) (first maximum wins), max value
// Here one kernel reads both NHWC heat-maps, emits the averaged map and the per-(b,j) arg-max.
// The average is computed as (a + f) * 0.5f exactly like the reference's fp32 torch expression, so
// arg-max indices are bit-exact given bit-identical heat-maps.
#include "common.cuh"
#include "kernels.h"

namespace fpd {
namespace {

struct Best {
  float v;
  int i;
};
// numpy semantics: NaN counts as the maximum (np.argmax returns the first NaN, np.amax NaN), otherwise the first
// occurrence of the largest value
__device__ __forceinline__ bool beats(float v, float best) { return v > best || (v != v && best == best); }
__device__ __forceinline__ Best better(Best a, Best b) {
  const bool an = a.v != a.v, bn = b.v != b.v;
  if (an || bn) {
    if (an && bn) return b.i < a.i ? b : a;
    return bn ? b : a;
  }
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}

// grid: B blocks; thread t -> joint j = t % J, pixel lane = t / J, stepping blockDim/J pixels.
__global__ void flip_merge_argmax_kernel(const float* __restrict__ hm, const float* __restrict__ hm_flip,
                                         const int* __restrict__ perm, int shift, float* __restrict__ avg_nhwc,
                                         int* __restrict__ idx, float* __restrict__ maxval, int J, int h, int w) {
  extern __shared__ unsigned char smraw[];
  Best* sbest = reinterpret_cast<Best*>(smraw);
  const int b = blockIdx.x;
  const int lanes = blockDim.x / J;
  const int j = threadIdx.x % J, pl = threadIdx.x / J;
  const int hw = h * w;
  const float* a = hm + (int64_t)b * hw * J;
  const float* f = hm_flip ? hm_flip + (int64_t)b * hw * J : nullptr;
  const int pj = (f && perm) ? perm[j] : j;
  Best best{-INFINITY, 0x7fffffff};
  if (pl < lanes) {
    for (int p = pl; p < hw; p += lanes) {
      float v = __ldg(a + (int64_t)p * J + j);
      if (f) {
        const int y = p / w, x = p % w;
        // flipped-back value at x comes from mirrored column; with the 1-px shift column x reads x-1 (x>=1)
        const int xs = (shift && x >= 1) ? x - 1 : x;
        const float fv = __ldg(f + ((int64_t)y * w + (w - 1 - xs)) * J + pj);
        v = (v + fv) * 0.5f;
      }
      if (avg_nhwc) avg_nhwc[((int64_t)b * hw + p) * J + j] = v;
      if (beats(v, best.v)) { best.v = v; best.i = p; }
    }
    sbest[pl * J + j] = best;
  }
  __syncthreads();
  if (threadIdx.x < J) {
    Best r = sbest[threadIdx.x];
    for (int l = 1; l < lanes; ++l) r = better(r, sbest[l * J + threadIdx.x]);
    idx[b * J + threadIdx.x] = r.i == 0x7fffffff ? 0 : r.i;
    maxval[b * J + threadIdx.x] = r.v;
  }
}

// plain NCHW arg-max: one warp per (b,j) map
__global__ void argmax_nchw_kernel(const float* __restrict__ hm, int* __restrict__ idx, float* __restrict__ maxval,
                                   int BJ, int hw) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= BJ) return;
  const float* p = hm + (int64_t)warp * hw;
  Best best{-INFINITY, 0x7fffffff};
  for (int i = lane; i < hw; i += 32) {
    const float v = __ldg(p + i);
    if (beats(v, best.v)) { best.v = v; best.i = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Best other;
    other.v = __shfl_xor_sync(0xffffffffu, best.v, o);
    other.i = __shfl_xor_sync(0xffffffffu, best.i, o);
    best = better(best, other);
  }
  if (lane == 0) {
    idx[warp] = best.i == 0x7fffffff ? 0 : best.i;
    maxval[warp] = best.v;
  }
}

}  // namespace

int flip_merge_argmax(const float* hm, const float* hm_flip, const int* flip_perm, int shift, float* avg_nhwc,
                      int* idx, float* maxval, int B, int J, int h, int w, cudaStream_t stream) {
  FPD_REQUIRE(J >= 1 && J <= 64, "flip_merge_argmax: J=%d out of range", J);
  const int threads = (512 / J) * J;
  const size_t smem = (size_t)threads * sizeof(Best);
  flip_merge_argmax_kernel<<<B, threads, smem, stream>>>(hm, hm_flip, flip_perm, shift, avg_nhwc, idx, maxval, J, h,
                                                         w);
  FPD_LAUNCH_CHECK();
  return FPD_OK;
}

int argmax_nchw(const float* hm, int* idx, float* maxval, int BJ, int hw, cudaStream_t stream) {
  const int threads = 256;
  const int blocks = (BJ * 32 + threads - 1) / threads;
  argmax_nchw_kernel<<<blocks, threads, 0, stream>>>(hm, idx, maxval, BJ, hw);
  FPD_LAUNCH_CHECK();
  return FPD_OK;
}

}  // namespace fpd
