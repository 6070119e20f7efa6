// conv_tc3.cu -- implicit-GEMM convolution, operand transform fused, A operand fed from TENSOR MEMORY (TS mode).
//
// Why: with both operands in shared memory (conv_tc.cu / conv_tc2.cu) a 128xNx8 tf32 MMA reads 4 KB of A and N*32 B of
// B from shared memory; at 128 B/clk/SM that is (4096 + 32 N)/128 cycles against N/2 cycles of tensor work, so for
// N <= 128 the tensor core is shared-memory-bandwidth bound (N = 64: 48 vs 32 cycles), and the 3xTF32 scheme reads the
// A tile three times. Here the transform warps write the prepared (a_hi, a_lo) tiles straight into TMEM with tcgen05.st
// and the MMAs take A from TMEM ("tcgen05.mma [d], [a_tmem], b_desc"): shared memory then carries only the raw x tile
// (TMA write + one read) and the weight tiles.
//   TMEM map (512 columns): [accumulator stage 0 | accumulator stage 1 | A stages: {hi 32 cols, lo 32 cols} x S]
//   A[m][k] lives at lane m, column k (one tf32 per 32-bit column), the layout UMMA expects for M = 128.
//
// [conv_tc2.cu header follows]
// conv_tc2.cu -- implicit-GEMM convolution with the operand transform fused into the pipeline.
//
// Same GEMM, tiling, TMA boxes and tcgen05 issue scheme as conv_tc.cu, but the A operand is read from HBM/L2 as the
// RAW fp32 activation x (one TMA box instead of a hi/lo pair), and eight "transform" warps turn each landed tile
//     x  ->  a = relu?((x - mean[c]) * scale[c] + shift[c])  ->  (a_hi, a_lo)   [tf32 split for 3xTF32]
// in place in shared memory before the MMA warp consumes it. That is BatchNorm-apply + ReLU + operand split of the
// reference's  conv(relu(bn(x)))  (lib/models/hourglass.py:34-44) without the separate HBM pass (read 4 B + write 8 B
// per element) and with half the A-operand traffic of the pre-split kernel. Padding positions of the 3x3 taps (and
// rows of a ragged last batch tile, channels of a ragged last k-block) are forced to zero AFTER the affine, exactly
// like Conv2d(padding=1) on the activated tensor.
//
// Pipeline per stage:  TMA (x tile, W_hi, W_lo) --full[s]--> transform warps --ready[s]--> MMA --empty[s]--> TMA
// Shared-memory accesses of the transform are bank-conflict free: thread (row r, half h) visits physical 16-byte chunk
// p = i ^ (r & 7), which under SWIZZLE_128B is logical chunk i for every row, so one iteration handles the same four
// channels in all threads (broadcast parameter reads) while the 8 threads of a quarter-warp hit 8 different chunks.
//
// Warp roles (448 threads, persistent): warp 0 TMA producer, warp 1 TMEM alloc + MMA issuer, warps 2-5 epilogue,
// warps 6-13 transform.
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace fpd {
namespace {

constexpr int kTileM = 128;
constexpr int kBlockK = 32;
constexpr int kABytes = kTileM * 128;
constexpr int kThreads = 448;
constexpr int kTransformThreads = 256;
constexpr int kMaxCin = 512;

struct Conv2Params {
  int B, H, W, Cin, Cout;
  int taps, passes;
  int bn, bh, bw;
  int tiles_w, tiles_h, tiles_n, num_tiles;
  int stages, b_bytes, tmem_cols, acc_stages, a_col0;
  int epi_prefetch;  // prefetch the residual row (L2 + one chunk ahead in registers) in the epilogue
  int nt, n_tiles;   // output channels are processed in n_tiles slices of nt (<= 256) columns; tile = m_tile * n_tiles + n_tile
  const float* pre_mean;   // [Cin] or null (0)
  const float* pre_scale;  // [Cin] or null (identity affine)
  const float* pre_shift;  // [Cin]
  int pre_relu;
  const float* bias;
  const float* residual;
  const float* relu_mask;
  float* y;
  float out_scale;
};

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
conv_tc_ts_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w_hi,
                     const __grid_constant__ CUtensorMap tm_w_lo, const Conv2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const bool split = p.passes == 3;
  // stage layout: [x tile 16K][B_hi][B_lo]  (B_lo only in 3xTF32 mode)
  const int stage_bytes = kABytes + (split ? 2 : 1) * p.b_bytes;
  uint8_t* tail = smem + (size_t)p.stages * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* ready_bar = full_bar + p.stages;
  uint64_t* empty_bar = ready_bar + p.stages;
  uint64_t* tmem_full = empty_bar + p.stages;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* s_mean = reinterpret_cast<float*>(tail + 256);  // [kMaxCin] x 3
  float* s_scale = s_mean + kMaxCin;
  float* s_shift = s_scale + kMaxCin;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w_hi);
    if (split) tma_prefetch_desc(&tm_w_lo);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&ready_bar[s], kTransformThreads / 32);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 128);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc_dyn(tmem_ptr_smem, 512u);
  const int cin_pad = (p.Cin + kBlockK - 1) / kBlockK * kBlockK;
  for (int c = threadIdx.x; c < cin_pad; c += kThreads) {
    const bool in = c < p.Cin;
    s_mean[c] = (in && p.pre_mean) ? p.pre_mean[c] : 0.f;
    s_scale[c] = (in && p.pre_scale) ? p.pre_scale[c] : 1.f;
    s_shift[c] = (in && p.pre_scale) ? p.pre_shift[c] : 0.f;
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int kblocks_per_tap = cin_pad / kBlockK;
  const int kblocks = p.taps * kblocks_per_tap;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const uint32_t tx_bytes = (uint32_t)(kABytes + (split ? 2 : 1) * p.b_bytes);
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int mt = tile / p.n_tiles;
        const int tw = mt % p.tiles_w;
        const int th = (mt / p.tiles_w) % p.tiles_h;
        const int tn = mt / (p.tiles_w * p.tiles_h);
        const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
        const int n0w = (tile % p.n_tiles) * p.nt;   // first output channel of this slice
        for (int tap = 0; tap < p.taps; ++tap) {
          const int dh = (p.taps == 9) ? (tap / 3 - 1) : 0;
          const int dw = (p.taps == 9) ? (tap % 3 - 1) : 0;
          for (int cb = 0; cb < kblocks_per_tap; ++cb, s = (s + 1 == p.stages ? 0 : s + 1), ph ^= (s == 0)) {
            mbar_wait(&empty_bar[s], ph ^ 1);
            uint8_t* st = smem + (size_t)s * stage_bytes;
            mbar_expect_tx(&full_bar[s], tx_bytes);
            tma_load_4d(st, &tm_x, &full_bar[s], cb * kBlockK, w0 + dw, h0 + dh, n0);
            tma_load_3d(st + kABytes, &tm_w_hi, &full_bar[s], cb * kBlockK, n0w, tap);
            if (split) tma_load_3d(st + kABytes + p.b_bytes, &tm_w_lo, &full_bar[s], cb * kBlockK, n0w, tap);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(kTileM, (uint32_t)p.nt, 0, 0);
      uint32_t tile_iter = 0;
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tile_iter) {
        const uint32_t as = p.acc_stages == 2 ? (tile_iter & 1) : 0u;
        const uint32_t aph = (p.acc_stages == 2 ? (tile_iter >> 1) : tile_iter) & 1;
        mbar_wait(&tmem_empty[as], aph ^ 1);
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + as * (uint32_t)p.tmem_cols;
        for (int kb = 0; kb < kblocks; ++kb, s = (s + 1 == p.stages ? 0 : s + 1), ph ^= (s == 0)) {
          mbar_wait(&ready_bar[s], ph);  // A tile is in TMEM (implies the TMA bytes of this stage have landed)
          tc_fence_after_sync();
          const uint32_t b_hi = smem_u32(smem + (size_t)s * stage_bytes) + kABytes;
          const uint32_t b_lo = b_hi + p.b_bytes;
          const uint32_t ta_hi = tmem_base + (uint32_t)p.a_col0 + (uint32_t)s * 64u;
          const uint32_t ta_lo = ta_hi + 32u;
#pragma unroll
          for (int ks = 0; ks < kBlockK / 8; ++ks) {
            const uint64_t db_hi = umma_desc_sw128(b_hi + ks * 32, 16, 1024);
            uint32_t acc = (kb > 0 || ks > 0) ? 1u : 0u;
            if (split) {
              const uint64_t db_lo = umma_desc_sw128(b_lo + ks * 32, 16, 1024);
              umma_tf32_ts(tmem_d, ta_lo + ks * 8, db_hi, idesc, acc);
              umma_tf32_ts(tmem_d, ta_hi + ks * 8, db_lo, idesc, 1u);
              acc = 1u;
            }
            umma_tf32_ts(tmem_d, ta_hi + ks * 8, db_hi, idesc, acc);
          }
          umma_commit(&empty_bar[s]);   // frees the smem stage AND the TMEM A stage
        }
        umma_commit(&tmem_full[as]);
      }
    }
  } else if (warp < 6) {
    // ===================== epilogue =====================
    const int q = warp & 3;
    const int m = q * 32 + lane;
    uint32_t tile_iter = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tile_iter) {
      const uint32_t as = p.acc_stages == 2 ? (tile_iter & 1) : 0u;
      const uint32_t aph = (p.acc_stages == 2 ? (tile_iter >> 1) : tile_iter) & 1;
      const int mt = tile / p.n_tiles;
        const int tw = mt % p.tiles_w;
      const int th = (mt / p.tiles_w) % p.tiles_h;
      const int tn = mt / (p.tiles_w * p.tiles_h);
      const int pw = tw * p.bw + (m % p.bw);
      const int ph_ = th * p.bh + (m / p.bw) % p.bh;
      const int pn = tn * p.bn + m / (p.bw * p.bh);
      const bool valid = pn < p.B;
      const size_t pix = ((size_t)pn * p.H + ph_) * p.W + pw;
      const int nc0 = (tile % p.n_tiles) * p.nt;   // channel slice of this tile
      float* yrow = p.y + pix * p.Cout + nc0;
      const float* rrow = p.residual ? p.residual + pix * p.Cout + nc0 : nullptr;
      const float* mrow = p.relu_mask ? p.relu_mask + pix * p.Cout + nc0 : nullptr;
      const float* brow = p.bias ? p.bias + nc0 : nullptr;
      // The residual row does not depend on the MMAs: pull it towards the SM while the accumulator is still being
      // produced (L2 prefetch of the whole row now, register prefetch one 16-column chunk ahead below); otherwise every
      // chunk would expose a full global-memory round trip and the epilogue, not the tensor core, sets the tile time.
      if (rrow && valid && p.epi_prefetch) {
        for (int c = 0; c < p.nt; c += 32)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(rrow + c));
      }
      mbar_wait(&tmem_full[as], aph);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + as * (uint32_t)p.tmem_cols + ((uint32_t)(q * 32) << 16);
      float4 rnext[4];
      if (rrow && valid && p.epi_prefetch) {
#pragma unroll
        for (int j = 0; j < 4; ++j) rnext[j] = __ldg(reinterpret_cast<const float4*>(rrow) + j);
      }
      for (int c0 = 0; c0 < p.nt; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + c0, v);
        float4 rcur[4];
        if (rrow && valid) {
          if (p.epi_prefetch) {
#pragma unroll
            for (int j = 0; j < 4; ++j) rcur[j] = rnext[j];
            if (c0 + 16 < p.nt) {
#pragma unroll
              for (int j = 0; j < 4; ++j) rnext[j] = __ldg(reinterpret_cast<const float4*>(rrow + c0 + 16) + j);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) rcur[j] = __ldg(reinterpret_cast<const float4*>(rrow + c0) + j);
          }
        }
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float4 o;
            o.x = __uint_as_float(v[j + 0]) * p.out_scale;
            o.y = __uint_as_float(v[j + 1]) * p.out_scale;
            o.z = __uint_as_float(v[j + 2]) * p.out_scale;
            o.w = __uint_as_float(v[j + 3]) * p.out_scale;
            if (brow) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(brow + c0 + j));
              o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
            }
            if (rrow) {
              const float4 r = rcur[j >> 2];
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
  } else {
    // ===================== operand transform -> TMEM =====================
    // warp (q = warp & 3, half = (warp - 6) >> 2): rows 32q..32q+31 (TMEM lane quarter q), channels 16*half..+15 of the
    // k-block. Thread = one pixel row: reads its 4 16-byte chunks from the swizzled x tile (conflict-free: the 8
    // threads of a quarter-warp touch 8 different physical chunks), applies the affine/ReLU, splits, and stores 16 hi
    // + 16 lo columns with tcgen05.st.
    const int q = warp & 3;
    const int half = (warp - 6) >> 2;
    const int r = q * 32 + lane;
    const int dn = r / (p.bw * p.bh), dh_ = (r / p.bw) % p.bh, dw_ = r % p.bw;
    const bool has_affine = p.pre_scale != nullptr;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)p.a_col0 + (uint32_t)half * 16u;
    int s = 0;
    uint32_t ph = 0;
    const uint32_t s_mean_a = smem_u32(s_mean), s_scale_a = smem_u32(s_scale), s_shift_a = smem_u32(s_shift);
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int mt = tile / p.n_tiles;
        const int tw = mt % p.tiles_w;
      const int th = (mt / p.tiles_w) % p.tiles_h;
      const int tn = mt / (p.tiles_w * p.tiles_h);
      const int pn = tn * p.bn + dn;
      const int h_base = th * p.bh + dh_, w_base = tw * p.bw + dw_;
      for (int tap = 0; tap < p.taps; ++tap) {
        const int hh = h_base + ((p.taps == 9) ? (tap / 3 - 1) : 0);
        const int ww = w_base + ((p.taps == 9) ? (tap % 3 - 1) : 0);
        const bool inb = pn < p.B && hh >= 0 && hh < p.H && ww >= 0 && ww < p.W;
        for (int cb = 0; cb < kblocks_per_tap; ++cb, s = (s + 1 == p.stages ? 0 : s + 1), ph ^= (s == 0)) {
          mbar_wait(&full_bar[s], ph);
          const uint32_t xrow = smem_u32(smem + (size_t)s * stage_bytes) + (uint32_t)r * 128u;
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int ii = 0; ii < 4; ++ii) {
            const int i = half * 4 + ii;
            const int c = cb * kBlockK + i * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (inb && c < p.Cin) {
              v = lds128(xrow + (uint32_t)((i ^ (r & 7)) << 4));
              if (has_affine) {
                const float4 mu = lds128(s_mean_a + (uint32_t)c * 4u);
                const float4 sc = lds128(s_scale_a + (uint32_t)c * 4u);
                const float4 sh = lds128(s_shift_a + (uint32_t)c * 4u);
                v.x = fmaf(v.x - mu.x, sc.x, sh.x); v.y = fmaf(v.y - mu.y, sc.y, sh.y);
                v.z = fmaf(v.z - mu.z, sc.z, sh.z); v.w = fmaf(v.w - mu.w, sc.w, sh.w);
              }
              if (p.pre_relu) {
                v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
              }
            }
            float4 h, l;
            split_tf32_fast(v.x, h.x, l.x); split_tf32_fast(v.y, h.y, l.y);
            split_tf32_fast(v.z, h.z, l.z); split_tf32_fast(v.w, h.w, l.w);
            hi[ii * 4 + 0] = __float_as_uint(h.x); hi[ii * 4 + 1] = __float_as_uint(h.y);
            hi[ii * 4 + 2] = __float_as_uint(h.z); hi[ii * 4 + 3] = __float_as_uint(h.w);
            lo[ii * 4 + 0] = __float_as_uint(l.x); lo[ii * 4 + 1] = __float_as_uint(l.y);
            lo[ii * 4 + 2] = __float_as_uint(l.z); lo[ii * 4 + 3] = __float_as_uint(l.w);
          }
          const uint32_t ta = lane_base + (uint32_t)s * 64u;
          tmem_st16(ta, hi);
          if (split) tmem_st16(ta + 32u, lo);
          tmem_st_wait();
          tc_fence_before_sync();
          __syncwarp();                   // one arrival per warp: 256 per-thread arrivals on one barrier serialise
          if (lane == 0) mbar_arrive(&ready_bar[s]);
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc_dyn(tmem_base, 512u);
  }
}

int pow2_floor_div(int x, int cap) {
  int r = 1;
  while (r * 2 <= cap && x % (r * 2) == 0) r *= 2;
  return r;
}

}  // namespace

// Output channels are processed in slices of nt columns: the widest divisor of Cout that is a multiple of 16 and at
// most 128, so that two accumulator stages (2 x 128 columns) and four A stages (4 x 64 columns) fit the 512 TMEM
// columns and the epilogue of one slice overlaps the MMAs of the next (Cout = 256 -> 2 x 128, 384 -> 3 x 128, 192 -> 2 x 96).
int conv_tc_ts_slice(int Cout) {
  if (Cout <= 128) return Cout;
  for (int nt = 128; nt >= 16; nt -= 16)
    if (Cout % nt == 0) return nt;
  return 0;
}

bool conv_tc_ts_supported(int Cin, int Cout, int ksize) {
  return (ksize == 1 || ksize == 3) && Cin % 4 == 0 && Cin >= 4 && Cin <= kMaxCin && Cout % 16 == 0 && Cout >= 16 &&
         Cout <= 1024 && conv_tc_ts_slice(Cout) > 0;
}

int conv_tc_ts_launch(const float* x, const float* pre_mean, const float* pre_scale, const float* pre_shift,
                         int pre_relu, const float* w_hi, const float* w_lo, const float* bias, const float* residual,
                         const float* relu_mask, float* y, float out_scale, int B, int H, int W, int Cin, int Cout,
                         int ksize, int num_sms, cudaStream_t stream) {
  FPD_REQUIRE(conv_tc_ts_supported(Cin, Cout, ksize), "conv_tc_ts: unsupported shape Cin=%d Cout=%d k=%d", Cin, Cout,
              ksize);
  FPD_REQUIRE(Cin <= kMaxCin, "conv_tc_ts: Cin=%d exceeds %d", Cin, kMaxCin);
  FPD_REQUIRE(x && w_hi && y, "conv_tc_ts: null operand");
  FPD_REQUIRE((pre_scale == nullptr) == (pre_shift == nullptr), "conv_tc_ts: pre_scale/pre_shift come in pairs");
  FPD_REQUIRE(pre_scale != nullptr || pre_mean == nullptr, "conv_tc_ts: pre_mean needs pre_scale/pre_shift");
  Conv2Params p{};
  p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.taps = ksize * ksize;
  p.passes = w_lo ? 3 : 1;
  p.bw = pow2_floor_div(W, kTileM);
  p.bh = pow2_floor_div(H, kTileM / p.bw);
  p.bn = kTileM / (p.bw * p.bh);
  p.tiles_w = W / p.bw; p.tiles_h = H / p.bh; p.tiles_n = (B + p.bn - 1) / p.bn;
  {
    static const int pf = [] { const char* e = getenv("FPD_EPI_PREFETCH"); return (e && e[0] == '1') ? 1 : 0; }();   // measured slightly slower (tools/bench_conv_variants.py): off
    p.epi_prefetch = pf;
  }
  p.nt = conv_tc_ts_slice(Cout);
  p.n_tiles = Cout / p.nt;
  p.num_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles;
  p.b_bytes = p.nt * 128;
  int tc = 32;
  while (tc < p.nt) tc *= 2;
  p.tmem_cols = tc;
  p.pre_mean = pre_mean; p.pre_scale = pre_scale; p.pre_shift = pre_shift; p.pre_relu = pre_relu;
  p.bias = bias; p.residual = residual; p.relu_mask = relu_mask; p.y = y; p.out_scale = out_scale;
  const int stage_bytes = kABytes + (p.passes == 3 ? 2 : 1) * p.b_bytes;
  const int tail_bytes = 256 + 3 * kMaxCin * (int)sizeof(float);
  int stages = (210 * 1024 - tail_bytes) / stage_bytes;
  // TMEM: accumulators (double-buffered when they fit) + 64 columns of A (hi, lo) per stage, 512 columns in all
  p.acc_stages = (2 * p.tmem_cols + 2 * 64 <= 512) ? 2 : 1;
  p.a_col0 = p.acc_stages * p.tmem_cols;
  const int tmem_stages = (512 - p.a_col0) / 64;
  if (stages > tmem_stages) stages = tmem_stages;
  if (stages > 6) stages = 6;
  FPD_REQUIRE(stages >= 2, "conv_tc_ts: tile does not fit in shared memory (stage=%d B)", stage_bytes);
  p.stages = stages;
  const size_t smem_bytes = (size_t)stages * stage_bytes + 1024 + tail_bytes;

  CUtensorMap tm_x, tm_w_hi, tm_w_lo;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t strides[3] = {(uint64_t)Cin * 4, (uint64_t)W * Cin * 4, (uint64_t)H * W * Cin * 4};
    uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
    int rc = encode_tmap(&tm_x, x, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, (uint64_t)p.taps};
    uint64_t strides[2] = {(uint64_t)Cin * 4, (uint64_t)Cout * Cin * 4};
    uint32_t box[3] = {(uint32_t)kBlockK, (uint32_t)p.nt, 1};
    int rc = encode_tmap(&tm_w_hi, w_hi, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = encode_tmap(&tm_w_lo, w_lo ? w_lo : w_hi, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    FPD_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  conv_tc_ts_kernel<<<grid, kThreads, smem_bytes, stream>>>(tm_x, tm_w_hi, tm_w_lo, p);
  FPD_LAUNCH_CHECK();
  return FPD_OK;
}

}  // namespace fpd
