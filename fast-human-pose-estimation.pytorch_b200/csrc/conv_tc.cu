// conv_tc.cu -- im2col-free implicit-GEMM convolution (1x1 / 3x3, stride 1, "same" padding) on the
// sm_100a tensor cores: TMA-staged NHWC tiles -> SWIZZLE_128B shared memory -> tcgen05.mma (kind::tf32)
// -> fp32 accumulators in TMEM -> tcgen05.ld epilogue (bias / residual / ReLU-mask fused).
//
// Replaces the cuDNN/oneDNN convolution the reference reaches through nn.Conv2d
// (reference lib/models/hourglass.py:20-27 conv1/conv2/conv3, :134-137,149,163 fc/score/fc_/score_).
//
// GEMM view:  D[pixel, co] = sum_{tap, ci} A[pixel (+) tap, ci] * Wt[tap][co][ci]
//   M = 128 pixels per tile (a bn x bh x bw box of the NHWC tensor), N = Cout (one UMMA, <= 256),
//   K = taps * Cin, consumed in k-blocks of 32 channels (= one 128-byte swizzle row of tf32).
// The 3x3 halo is never materialised: each tap is a TMA box load at (h0+dh, w0+dw) and the TMA unit
// zero-fills out-of-bounds rows/columns (the padding of the reference's Conv2d(padding=1)).
//
// Precision: `passes == 3` runs the 3xTF32 scheme (a_hi*b_hi + a_lo*b_hi + a_hi*b_lo, fp32 accumulate)
// which is what lets the network meet the 1e-3 parity bar against the fp32 reference;
// `passes == 1` is plain TF32.
//
// Warp roles (192 threads, persistent over tiles):
//   warp 0      : TMA producer (one elected lane)
//   warp 1      : TMEM allocator + MMA issuer (one elected lane)
//   warps 2..5  : epilogue; warp w owns TMEM lanes [32*(w%4), 32*(w%4)+32)
#include "common.cuh"
#include "kernels.h"

namespace fpd {

namespace {

constexpr int kTileM = 128;
constexpr int kBlockK = 32;               // tf32 elements per 128B swizzle row
constexpr int kABytes = kTileM * 128;     // one A tile (hi or lo): 128 rows x 128 B
constexpr int kThreads = 192;

struct ConvTcParams {
  int B, H, W, Cin, Cout;
  int taps;        // 1 or 9
  int passes;      // 1 or 3
  int bn, bh, bw;  // pixel-tile box
  int tiles_w, tiles_h, tiles_n, num_tiles;
  int stages;
  int b_bytes;     // bytes of one B tile (hi or lo) = Cout * 128
  int tmem_cols;   // per accumulator stage (pow2 >= 32)
  const float* bias;       // [Cout] or null
  const float* residual;   // NHWC [B,H,W,Cout] or null
  const float* relu_mask;  // NHWC [B,H,W,Cout] or null : y = mask > 0 ? acc : 0
  float* y;                // NHWC [B,H,W,Cout]
  float out_scale;         // y = acc * out_scale (+bias +residual)
};

__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
               const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
               const ConvTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operands need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int stage_bytes = (p.passes == 3 ? 2 : 1) * (kABytes + p.b_bytes);
  uint8_t* bar_region = smem + (size_t)p.stages * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_region);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full = empty_bar + p.stages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;         // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a_hi);
    tma_prefetch_desc(&tm_w_hi);
    if (p.passes == 3) {
      tma_prefetch_desc(&tm_a_lo);
      tma_prefetch_desc(&tm_w_lo);
    }
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 128);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc_dyn(tmem_ptr_smem, (uint32_t)(2 * p.tmem_cols));
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int kblocks_per_tap = (p.Cin + kBlockK - 1) / kBlockK;  // a ragged last block is zero-filled by TMA
  const int kblocks = p.taps * kblocks_per_tap;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int tw = tile % p.tiles_w;
        const int th = (tile / p.tiles_w) % p.tiles_h;
        const int tn = tile / (p.tiles_w * p.tiles_h);
        const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int dh = (p.taps == 9) ? (tap / 3 - 1) : 0;
          const int dw = (p.taps == 9) ? (tap % 3 - 1) : 0;
          for (int cb = 0; cb < kblocks_per_tap; ++cb, s = (s + 1 == p.stages ? 0 : s + 1), ph ^= (s == 0)) {
            mbar_wait(&empty_bar[s], ph ^ 1);
            uint8_t* st = smem + (size_t)s * stage_bytes;
            mbar_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
            tma_load_4d(st, &tm_a_hi, &full_bar[s], cb * kBlockK, w0 + dw, h0 + dh, n0);
            tma_load_3d(st + kABytes, &tm_w_hi, &full_bar[s], cb * kBlockK, 0, tap);
            if (p.passes == 3) {
              uint8_t* lo = st + kABytes + p.b_bytes;
              tma_load_4d(lo, &tm_a_lo, &full_bar[s], cb * kBlockK, w0 + dw, h0 + dh, n0);
              tma_load_3d(lo + kABytes, &tm_w_lo, &full_bar[s], cb * kBlockK, 0, tap);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(kTileM, (uint32_t)p.Cout, 0, 0);
      int s = 0;
      uint32_t ph = 0;
      uint32_t tile_iter = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tile_iter) {
        const uint32_t as = tile_iter & 1;
        const uint32_t aph = (tile_iter >> 1) & 1;
        mbar_wait(&tmem_empty[as], aph ^ 1);
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + as * (uint32_t)p.tmem_cols;
        for (int kb = 0; kb < kblocks; ++kb, s = (s + 1 == p.stages ? 0 : s + 1), ph ^= (s == 0)) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after_sync();
          const uint32_t a_hi = smem_u32(smem + (size_t)s * stage_bytes);
          const uint32_t b_hi = a_hi + kABytes;
          const uint32_t a_lo = b_hi + p.b_bytes;
          const uint32_t b_lo = a_lo + kABytes;
#pragma unroll
          for (int ks = 0; ks < kBlockK / 8; ++ks) {
            const uint32_t koff = ks * 32;  // 8 tf32 = 32 bytes along K inside the swizzle row
            const uint64_t da_hi = umma_desc_sw128(a_hi + koff, 16, 1024);
            const uint64_t db_hi = umma_desc_sw128(b_hi + koff, 16, 1024);
            uint32_t acc = (kb > 0 || ks > 0) ? 1u : 0u;
            if (p.passes == 3) {
              const uint64_t da_lo = umma_desc_sw128(a_lo + koff, 16, 1024);
              const uint64_t db_lo = umma_desc_sw128(b_lo + koff, 16, 1024);
              umma_tf32(tmem_d, da_lo, db_hi, idesc, acc);  // small terms first
              umma_tf32(tmem_d, da_hi, db_lo, idesc, 1u);
              acc = 1u;
            }
            umma_tf32(tmem_d, da_hi, db_hi, idesc, acc);
          }
          umma_commit(&empty_bar[s]);  // frees the smem stage once the MMAs above have read it
        }
        umma_commit(&tmem_full[as]);   // accumulator for this tile is complete
      }
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;         // row of the tile = pixel
    uint32_t tile_iter = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tile_iter) {
      const uint32_t as = tile_iter & 1;
      const uint32_t aph = (tile_iter >> 1) & 1;
      const int tw = tile % p.tiles_w;
      const int th = (tile / p.tiles_w) % p.tiles_h;
      const int tn = tile / (p.tiles_w * p.tiles_h);
      const int pw = tw * p.bw + (m % p.bw);
      const int ph_ = th * p.bh + (m / p.bw) % p.bh;
      const int pn = tn * p.bn + m / (p.bw * p.bh);
      const bool valid = pn < p.B;
      const size_t pix = ((size_t)pn * p.H + ph_) * p.W + pw;
      float* yrow = p.y + pix * p.Cout;
      const float* rrow = p.residual ? p.residual + pix * p.Cout : nullptr;
      const float* mrow = p.relu_mask ? p.relu_mask + pix * p.Cout : nullptr;

      mbar_wait(&tmem_full[as], aph);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + as * (uint32_t)p.tmem_cols + ((uint32_t)(q * 32) << 16);
      for (int c0 = 0; c0 < p.Cout; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + c0, v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float4 o;
            o.x = __uint_as_float(v[j + 0]) * p.out_scale;
            o.y = __uint_as_float(v[j + 1]) * p.out_scale;
            o.z = __uint_as_float(v[j + 2]) * p.out_scale;
            o.w = __uint_as_float(v[j + 3]) * p.out_scale;
            if (p.bias) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + j));
              o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
            }
            if (rrow) {
              const float4 r = __ldg(reinterpret_cast<const float4*>(rrow + c0 + j));
              o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
            }
            if (mrow) {
              const float4 k = __ldg(reinterpret_cast<const float4*>(mrow + c0 + j));
              o.x = k.x > 0.f ? o.x : 0.f;
              o.y = k.y > 0.f ? o.y : 0.f;
              o.z = k.z > 0.f ? o.z : 0.f;
              o.w = k.w > 0.f ? o.w : 0.f;
            }
            *reinterpret_cast<float4*>(yrow + c0 + j) = o;
          }
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&tmem_empty[as]);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc_dyn(tmem_base, (uint32_t)(2 * p.tmem_cols));
  }
}

int pow2_floor_div(int x, int cap) {  // largest power of two dividing x, capped
  int r = 1;
  while (r * 2 <= cap && x % (r * 2) == 0) r *= 2;
  return r;
}

}  // namespace

bool conv_tc_supported(int Cin, int Cout, int ksize) {
  // Cin need not be a multiple of the 32-channel k-block: the TMA unit zero-fills the out-of-bounds channels of
  // both operands (16-channel score_ convs, HRNet-w48's 48/96-channel branches)
  return (ksize == 1 || ksize == 3) && Cin % 4 == 0 && Cin >= 4 && Cout % 16 == 0 && Cout >= 16 && Cout <= 256;
}

int conv_tc_launch(const float* a_hi, const float* a_lo, const float* w_hi, const float* w_lo, const float* bias,
                   const float* residual, const float* relu_mask, float* y, float out_scale, int B, int H, int W,
                   int Cin, int Cout, int ksize, int num_sms, cudaStream_t stream) {
  FPD_REQUIRE(conv_tc_supported(Cin, Cout, ksize), "conv_tc: unsupported shape Cin=%d Cout=%d k=%d", Cin, Cout,
              ksize);
  FPD_REQUIRE(a_hi && w_hi && y, "conv_tc: null operand");
  FPD_REQUIRE((a_lo == nullptr) == (w_lo == nullptr), "conv_tc: a_lo and w_lo must both be given (3xTF32) or both null");
  ConvTcParams p{};
  p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.taps = ksize * ksize;
  p.passes = a_lo ? 3 : 1;
  p.bw = pow2_floor_div(W, kTileM);
  p.bh = pow2_floor_div(H, kTileM / p.bw);
  p.bn = kTileM / (p.bw * p.bh);
  p.tiles_w = W / p.bw;
  p.tiles_h = H / p.bh;
  p.tiles_n = (B + p.bn - 1) / p.bn;
  p.num_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  p.b_bytes = Cout * 128;
  int tc = 32;
  while (tc < Cout) tc *= 2;
  p.tmem_cols = tc;
  p.bias = bias; p.residual = residual; p.relu_mask = relu_mask; p.y = y; p.out_scale = out_scale;

  const int stage_bytes = (p.passes == 3 ? 2 : 1) * (kABytes + p.b_bytes);
  const int budget = 200 * 1024;
  int stages = budget / stage_bytes;
  if (stages > 8) stages = 8;
  FPD_REQUIRE(stages >= 2, "conv_tc: tile does not fit in shared memory (stage=%d B)", stage_bytes);
  p.stages = stages;
  const size_t smem_bytes = (size_t)stages * stage_bytes + 1024 /*align slack*/ + 256 /*barriers*/;

  CUtensorMap tm_a_hi, tm_a_lo, tm_w_hi, tm_w_lo;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t strides[3] = {(uint64_t)Cin * 4, (uint64_t)W * Cin * 4, (uint64_t)H * W * Cin * 4};
    uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
    int rc = encode_tmap(&tm_a_hi, a_hi, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = encode_tmap(&tm_a_lo, a_lo ? a_lo : a_hi, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, (uint64_t)p.taps};
    uint64_t strides[2] = {(uint64_t)Cin * 4, (uint64_t)Cout * Cin * 4};
    uint32_t box[3] = {(uint32_t)kBlockK, (uint32_t)Cout, 1};
    int rc = encode_tmap(&tm_w_hi, w_hi, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = encode_tmap(&tm_w_lo, w_lo ? w_lo : w_hi, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }

  static bool attr_set = false;
  if (!attr_set) {
    FPD_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  conv_tc_kernel<<<grid, kThreads, smem_bytes, stream>>>(tm_a_hi, tm_a_lo, tm_w_hi, tm_w_lo, p);
  FPD_LAUNCH_CHECK();
  return FPD_OK;
}

}  // namespace fpd
