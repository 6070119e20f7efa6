// nms.cu -- greedy box NMS for sm_100a, entirely on the device.
//
// What it replaces: reference lib/nms/nms_kernel.cu:33-77 (pairwise-IoU bit-mask kernel) and the host
// driver `_nms` :90-143 (cudaMalloc/cudaFree per call, full-mask D2H, serial host sweep).
// Same arithmetic contract: boxes pre-sorted by score, IoU with the "+1 pixel" convention, a box is
// suppressed when IoU > thresh (strict), result = indices (into the sorted list) that survive.
//
// B200 design:
//  * pass 1 (HBM/L2-bound, embarrassingly parallel): only the upper triangle of 64x64 tiles is launched
//    (the reference computes both triangles and throws the lower away); box tiles are staged as SoA in
//    shared memory so the inner loop is bank-conflict free; each thread owns one row box and emits one
//    uint64 suppression word per column tile.
//  * pass 2 (latency-bound, serial by nature): one CTA sweeps the mask on the device, 64 boxes at a
//    time: the 64 diagonal words are staged to shared memory, one thread resolves the chunk with register
//    bit-ops, then the whole CTA ORs the kept rows into the running "removed" bitmap (coalesced reads).
//    Only the keep list (<= n ints) and its length ever cross PCIe.
#include "common.cuh"
#include "kernels.h"

namespace fpd {
namespace {

constexpr int kT = 64;

__device__ __forceinline__ float iou_p1(float ax1, float ay1, float ax2, float ay2, float bx1, float by1, float bx2,
                                        float by2) {
  const float left = fmaxf(ax1, bx1), right = fminf(ax2, bx2);
  const float top = fmaxf(ay1, by1), bottom = fminf(ay2, by2);
  const float width = fmaxf(right - left + 1.f, 0.f), height = fmaxf(bottom - top + 1.f, 0.f);
  const float inter = width * height;
  const float sa = (ax2 - ax1 + 1.f) * (ay2 - ay1 + 1.f);
  const float sb = (bx2 - bx1 + 1.f) * (by2 - by1 + 1.f);
  return inter / (sa + sb - inter);
}

// linear block id -> (row_tile, col_tile) with col_tile >= row_tile
__global__ void __launch_bounds__(kT)
nms_mask_kernel(const float* __restrict__ boxes, int n, int box_dim, float thresh, int col_blocks,
                unsigned long long* __restrict__ mask) {
  // decode triangular index
  int t = blockIdx.x;
  int row = 0;
  int rem = col_blocks;
  while (t >= rem) { t -= rem; ++row; --rem; }
  const int col = row + t;

  __shared__ float cx1[kT], cy1[kT], cx2[kT], cy2[kT];
  const int cj = col * kT + threadIdx.x;
  if (cj < n) {
    const float* b = boxes + (size_t)cj * box_dim;
    cx1[threadIdx.x] = b[0]; cy1[threadIdx.x] = b[1]; cx2[threadIdx.x] = b[2]; cy2[threadIdx.x] = b[3];
  }
  __syncthreads();
  const int ri = row * kT + threadIdx.x;
  if (ri >= n) return;
  const float* a = boxes + (size_t)ri * box_dim;
  const float ax1 = a[0], ay1 = a[1], ax2 = a[2], ay2 = a[3];
  const int col_size = min(kT, n - col * kT);
  unsigned long long bits = 0ull;
  const int start = (row == col) ? threadIdx.x + 1 : 0;
  for (int i = start; i < col_size; ++i) {
    if (iou_p1(ax1, ay1, ax2, ay2, cx1[i], cy1[i], cx2[i], cy2[i]) > thresh) bits |= 1ull << i;
  }
  mask[(size_t)ri * col_blocks + col] = bits;
}

__global__ void __launch_bounds__(1024)
nms_sweep_kernel(const unsigned long long* __restrict__ mask, int n, int col_blocks, int* __restrict__ keep,
                 int* __restrict__ num_keep) {
  extern __shared__ unsigned long long sm[];
  unsigned long long* remv = sm;               // [col_blocks]
  unsigned long long* diag = sm + col_blocks;  // [64]
  __shared__ unsigned long long kept_bits;
  __shared__ int count;
  for (int j = threadIdx.x; j < col_blocks; j += blockDim.x) remv[j] = 0ull;
  if (threadIdx.x == 0) count = 0;
  __syncthreads();
  for (int c = 0; c < col_blocks; ++c) {
    const int base = c * kT;
    const int csize = min(kT, n - base);
    if (threadIdx.x < kT) diag[threadIdx.x] = threadIdx.x < csize ? mask[(size_t)(base + threadIdx.x) * col_blocks + c] : 0ull;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long cur = remv[c];
      unsigned long long kb = 0ull;
      int cnt = count;
      for (int b = 0; b < csize; ++b) {
        if (!((cur >> b) & 1ull)) {
          keep[cnt++] = base + b;
          kb |= 1ull << b;
          cur |= diag[b];
        }
      }
      count = cnt;
      kept_bits = kb;
    }
    __syncthreads();
    const unsigned long long kb = kept_bits;
    for (int j = c + 1 + threadIdx.x; j < col_blocks; j += blockDim.x) {
      unsigned long long acc = remv[j];
      unsigned long long bits = kb;
      while (bits) {
        const int b = __ffsll((long long)bits) - 1;
        bits &= bits - 1;
        acc |= mask[(size_t)(base + b) * col_blocks + j];
      }
      remv[j] = acc;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *num_keep = count;
}

}  // namespace

size_t nms_workspace_bytes(int n) {
  const size_t col_blocks = (size_t)(n + kT - 1) / kT;
  return (size_t)n * col_blocks * sizeof(unsigned long long) + 256;
}

int nms_device(const float* boxes_sorted_dev, int n, int box_dim, float thresh, int* keep_dev, int* num_keep_dev,
               void* workspace, size_t ws_bytes, cudaStream_t stream) {
  FPD_REQUIRE(box_dim >= 4, "nms: box_dim=%d must be >= 4 (x1,y1,x2,y2[,score])", box_dim);
  if (n <= 0) {
    FPD_CUDA_CHECK(cudaMemsetAsync(num_keep_dev, 0, sizeof(int), stream));
    return FPD_OK;
  }
  FPD_REQUIRE(ws_bytes >= nms_workspace_bytes(n), "nms: workspace too small");
  const int col_blocks = (n + kT - 1) / kT;
  FPD_REQUIRE(col_blocks <= 4096, "nms: n=%d too large", n);
  unsigned long long* mask = (unsigned long long*)workspace;
  const int tri = col_blocks * (col_blocks + 1) / 2;
  nms_mask_kernel<<<tri, kT, 0, stream>>>(boxes_sorted_dev, n, box_dim, thresh, col_blocks, mask);
  FPD_LAUNCH_CHECK();
  const size_t smem = (size_t)(col_blocks + kT) * sizeof(unsigned long long);
  nms_sweep_kernel<<<1, 1024, smem, stream>>>(mask, n, col_blocks, keep_dev, num_keep_dev);
  FPD_LAUNCH_CHECK();
  return FPD_OK;
}

}  // namespace fpd
