// wgrad_tc3.cu -- 3x3 convolution weight gradient, generation 3: one halo tile per pixel block, all taps per CTA.
//
//   dW[co][ci][tap] = scale * sum_{pixels p} dY[p][co] * A[p (+) tap][ci],   A = relu?((x - mean) * scale + shift)
//
// What bounded wgrad_tc2.cu (tools/diag_wgrad.py, 3x3 64->64 @64x64, B=32: 287 us against ~30 us of tensor work): a
// pipeline stage carried only 32 pixels x 2 taps (384 cycles of MMA) but cost ~0.5 us of barrier round trips, the nine
// taps were fetched and transformed as nine separately shifted TMA boxes, and dY was re-transformed by each of the five
// tap-group CTAs. Here a stage is still 32 pixels (one K block), but
//   * x arrives ONCE per stage as a (bw+2) x (bh+2) halo tile (one TMA box per 32-channel block, hardware zero fill) and
//     is transformed + split ONCE; a tap is then just a different START ROW of the same MN-major shared-memory tile:
//     TMA (SWIZZLE_128B_ATOM_32B) and UMMA (SWIZZLE_128B_BASE32B) both key the swizzle on absolute shared-memory address
//     bits, so an operand may start at any 128-byte row (measured: tools/diag_wgrad_shift.py, bit-identical results for
//     tiles shifted 1..3 rows off the atom boundary);
//   * one CTA accumulates up to 512/Cout taps (5 + 4 for Cout = 64) in tensor memory, so dY is transformed once or twice
//     per pixel instead of five times and a stage carries 5 x 12 MMAs (~1900 cycles) per barrier round trip.
// GEMM per tap: D[ci][co] += A_tap^T[ci][32 px] * dY[32 px][co], both operands MN-major (pixel rows of 128 bytes =
// 32 channels), M = Cin (64 or 128), N = Cout, K = 8 pixels per MMA, 3xTF32 (hi/lo pairs, operands split in place).
// Cin = 32 ("pair" mode): UMMA has no M = 32, so TWO taps share one M = 64 instruction -- the descriptor's leading-
// dimension byte offset, normally the distance to the next 32-channel block, is set to the distance between the two
// taps' start rows in the SAME halo block (128 B for a +1 step along w), which the address-keyed swizzle allows just as
// it allows arbitrary start rows. Pairs (0,1) (2,3) (4,5) (6,7) (7,8): all nine taps accumulate in one CTA
// (5 x Cout TMEM columns); rows 0..31 of the last pair repeat tap 7 and are dropped.
//
// Warp roles (448 threads): warp 0 TMA producer, warp 1 TMEM alloc + MMA issuer, warps 2-5 epilogue (after the K loop),
// warps 6-13 operand transform. Split-K over pixel blocks across CTAs; wgrad_reduce3_kernel sums the partials in a fixed
// order straight into the OIHW gradient (deterministic).
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace fpd {
namespace {

constexpr int kKp = 32;          // pixels per K block / pipeline stage
constexpr int kThreads = 448;
constexpr int kXf = 256;
constexpr int kMaxCin = 128;
constexpr int kTail = 256 + 3 * kMaxCin * 4;

struct Wgrad3Params {
  int B, H, W, Cin, Cout, passes;
  int bw, bh, tiles_w, tiles_h, num_ktiles;
  int halo_w, halo_h, halo_px;
  int cblks, nblk;          // 32-channel blocks of x (M side) and dY (N side)
  int xblk_bytes;           // bytes of one x halo block (rounded up to 512)
  int half_bytes;           // hi (or lo) part of a stage
  int stage_bytes, stages;
  int groups, tg;           // tap groups (CTAs per K range) and taps per group
  int splits, kt_per_split;
  int tmem_cols;
  int lane_map;             // accumulator row -> TMEM lane mapping for M = 64 (see acc_lane)
  int pair;                 // Cin = 32: two taps share one M = 64 MMA (see the MMA issuer); 0 otherwise
  float* partial;           // [splits][9][Cin][Cout]
  const float* pre_mean;
  const float* pre_scale;
  const float* pre_shift;
  int pre_relu;
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

// D[tmem] (+)= A[smem desc] * B[smem desc], tf32, issued by one elected lane of a converged warp
__device__ __forceinline__ void umma_tf32_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// TMEM lane of accumulator row m. M = 128: lane m. M = 64: the 64 rows sit in 16 lanes of each 32-lane quarter
// (lane_map 1, the sm_100 data-path layout for M = 64) -- 0 / 2 are the alternatives probed by tools/diag_wgrad3.py.
__device__ __forceinline__ int acc_row_of_lane(int lane128, int M, int lane_map) {
  if (M == 128) return lane128;
  if (lane_map == 0) return lane128 < 64 ? lane128 : -1;
  if (lane_map == 1) return (lane128 & 31) < 16 ? (lane128 >> 5) * 16 + (lane128 & 15) : -1;
  return (lane128 & 63) < 32 ? (lane128 >> 6) * 32 + (lane128 & 31) : -1;
}

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc3_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_g,
                 const Wgrad3Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tail = smem + (size_t)p.stages * p.stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* ready_bar = full_bar + 8;
  uint64_t* empty_bar = ready_bar + 8;
  uint64_t* done_bar = empty_bar + 8;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(done_bar + 1);
  float* s_mean = reinterpret_cast<float*>(tail + 256);
  float* s_scale = s_mean + kMaxCin;
  float* s_shift = s_scale + kMaxCin;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int g = blockIdx.x % p.groups;          // tap group
  const int sp = blockIdx.x / p.groups;         // K split
  const int tap0 = p.pair ? 0 : g * p.tg;
  const int ntaps = p.pair ? 5 : min(p.tg, 9 - tap0);      // pair mode: five tap PAIRS
  const int kt0 = sp * p.kt_per_split;
  const int kt1 = min(kt0 + p.kt_per_split, p.num_ktiles);
  const int nkt = max(kt1 - kt0, 0);
  const bool split = p.passes == 3;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_g);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&ready_bar[s], kXf / 32);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(done_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc_dyn(tmem_ptr_smem, (uint32_t)p.tmem_cols);
  for (int c = threadIdx.x; c < p.Cin; c += kThreads) {
    s_mean[c] = p.pre_mean ? p.pre_mean[c] : 0.f;
    s_scale[c] = p.pre_scale ? p.pre_scale[c] : 1.f;
    s_shift[c] = p.pre_scale ? p.pre_shift[c] : 0.f;
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t x_bytes = (uint32_t)(p.cblks * p.xblk_bytes);   // x part of a stage half; dY boxes follow

  if (warp == 0) {
    // ===================== TMA producer =====================
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < nkt; ++i, s = (s + 1 == p.stages ? 0 : s + 1), ph ^= (s == 0)) {
      mbar_wait(&empty_bar[s], ph ^ 1);
      if (elect_one()) {
        const int kt = kt0 + i;
        const int tw = kt % p.tiles_w, th = (kt / p.tiles_w) % p.tiles_h, n0 = kt / (p.tiles_w * p.tiles_h);
        const int w0 = tw * p.bw, h0 = th * p.bh;
        uint8_t* st = smem + (size_t)s * p.stage_bytes;
        mbar_expect_tx(&full_bar[s], (uint32_t)(p.cblks * p.halo_px * 128 + p.nblk * kKp * 128));
        for (int cb = 0; cb < p.cblks; ++cb)
          tma_load_4d(st + (size_t)cb * p.xblk_bytes, &tm_x, &full_bar[s], cb * 32, w0 - 1, h0 - 1, n0);
        for (int j = 0; j < p.nblk; ++j)
          tma_load_4d(st + x_bytes + (size_t)j * (kKp * 128), &tm_g, &full_bar[s], j * 32, w0, h0, n0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = umma_idesc_tf32((uint32_t)(p.pair ? 64 : p.Cin), (uint32_t)p.Cout, 1, 1);   // both MN-major
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < nkt; ++i, s = (s + 1 == p.stages ? 0 : s + 1), ph ^= (s == 0)) {
      mbar_wait(&ready_bar[s], ph);
      tc_fence_after_sync();
      if (elect_one()) {
        const uint32_t x_hi = smem_u32(smem + (size_t)s * p.stage_bytes);
        const uint32_t g_hi = x_hi + x_bytes;
        const uint32_t x_lo = x_hi + (uint32_t)p.half_bytes;
        const uint32_t g_lo = g_hi + (uint32_t)p.half_bytes;
        for (int t = 0; t < ntaps; ++t) {
          const int tap = p.pair ? (t < 4 ? 2 * t : 7) : tap0 + t;      // pair mode: first tap of the pair
          const int dh = tap / 3 - 1, dw = tap % 3 - 1;
          // leading-dimension offset of the A operand: next 32-channel block, or (pair mode) the second tap's start row
          uint32_t lbo = (uint32_t)p.xblk_bytes;
          if (p.pair) {
            const int tb = tap + 1, dh2 = tb / 3 - 1, dw2 = tb % 3 - 1;
            lbo = (uint32_t)(((dh2 - dh) * p.halo_w + (dw2 - dw)) * 128);
          }
          const uint32_t tmem_d = tmem_base + (uint32_t)(t * p.Cout);
#pragma unroll
          for (int ks = 0; ks < kKp / 8; ++ks) {
            // the 8 pixels of this K step are consecutive along w: halo row of the first one, shifted by the tap
            const int prow = (ks * 8) / p.bw, pcol = (ks * 8) % p.bw;
            const uint32_t xoff = (uint32_t)(((prow + 1 + dh) * p.halo_w + pcol + 1 + dw) * 128);
            const uint64_t da_hi = umma_desc_sw128_32b(x_hi + xoff, lbo, 512);
            const uint64_t db_hi = umma_desc_sw128_32b(g_hi + ks * 1024, kKp * 128, 512);
            uint32_t acc = (i > 0 || ks > 0) ? 1u : 0u;
            if (split) {
              const uint64_t da_lo = umma_desc_sw128_32b(x_lo + xoff, lbo, 512);
              const uint64_t db_lo = umma_desc_sw128_32b(g_lo + ks * 1024, kKp * 128, 512);
              umma_tf32_ss(tmem_d, da_lo, db_hi, idesc, acc);
              umma_tf32_ss(tmem_d, da_hi, db_lo, idesc, 1u);
              acc = 1u;
            }
            umma_tf32_ss(tmem_d, da_hi, db_hi, idesc, acc);
          }
        }
        umma_commit(&empty_bar[s]);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(done_bar);
    __syncwarp();
  } else if (warp < 6) {
    // ===================== epilogue: accumulators -> fp32 partials [split][tap][ci][co] =====================
    const int q = warp & 3;
    const int M = p.pair ? 64 : p.Cin;
    const int row = acc_row_of_lane(q * 32 + lane, M, p.lane_map);
    if (nkt > 0) {
      mbar_wait(done_bar, 0);
      tc_fence_after_sync();
    }
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int t = 0; t < ntaps; ++t) {
      // pair mode: accumulator row = (which tap of the pair) * 32 + input channel; the last pair's first half repeats tap 7
      int tap_of_row = tap0 + t, ci_of_row = row < 0 ? 0 : row;
      bool keep = row >= 0 && row < M;
      if (p.pair) {
        const int half = ci_of_row >> 5;
        tap_of_row = (t < 4 ? 2 * t : 7) + half;
        ci_of_row &= 31;
        if (t == 4 && half == 0) keep = false;
      }
      float* prow = p.partial + (((size_t)sp * 9 + tap_of_row) * p.Cin + ci_of_row) * p.Cout;
      for (int c0 = 0; c0 < p.Cout; c0 += 16) {
        uint32_t v[16];
        if (nkt > 0) {
          tmem_ld16(taddr + (uint32_t)(t * p.Cout + c0), v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0u;
        }
        if (keep) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(prow + c0 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                    __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        }
      }
    }
  } else {
    // ===================== operand transform (in place: raw fp32 -> tf32 hi, lo at + half_bytes) =====================
    // thread t: row r = t / 8 (+32, +64, +96) of every block, physical 16-byte chunk pc = t % 8 of that row. Under the
    // 32-byte-atom swizzle the 32-byte chunk index is XORed with (absolute row & 3).
    const int t = threadIdx.x - 6 * 32;
    const int r = t >> 3, pc = t & 7;
    const bool has_affine = p.pre_scale != nullptr;
    const uint32_t mean_a = smem_u32(s_mean), scale_a = smem_u32(s_scale), shift_a = smem_u32(s_shift);
    // halo coordinates of this thread's (up to four) x rows: constant over the K loop
    int hy[4], hx[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int R = r + 32 * u;
      hy[u] = R / p.halo_w;
      hx[u] = R - hy[u] * p.halo_w;
    }
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < nkt; ++i, s = (s + 1 == p.stages ? 0 : s + 1), ph ^= (s == 0)) {
      const int kt = kt0 + i;
      const int tw = kt % p.tiles_w, th = (kt / p.tiles_w) % p.tiles_h;
      const int w0 = tw * p.bw, h0 = th * p.bh;
      mbar_wait(&full_bar[s], ph);
      const uint32_t base = smem_u32(smem + (size_t)s * p.stage_bytes);
      // ---- x halo blocks: affine + ReLU, zero outside the image (conv padding), split
      for (int cb = 0; cb < p.cblks; ++cb) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int R = r + 32 * u;
          if (R >= p.halo_px) break;
          const uint32_t addr = base + (uint32_t)(cb * p.xblk_bytes) + (uint32_t)R * 128u + (uint32_t)pc * 16u;
          float4 v = lds128(addr);
          const int hh = h0 - 1 + hy[u], ww = w0 - 1 + hx[u];
          if (hh >= 0 && hh < p.H && ww >= 0 && ww < p.W) {
            const int lc = ((((pc >> 1) ^ (int)((addr >> 7) & 3u)) << 1) | (pc & 1));   // logical 16-byte chunk
            const uint32_t c = (uint32_t)(cb * 32 + lc * 4);
            if (has_affine) {
              const float4 mu = lds128(mean_a + c * 4u);
              const float4 sc = lds128(scale_a + c * 4u);
              const float4 sh = lds128(shift_a + c * 4u);
              v.x = fmaf(v.x - mu.x, sc.x, sh.x); v.y = fmaf(v.y - mu.y, sc.y, sh.y);
              v.z = fmaf(v.z - mu.z, sc.z, sh.z); v.w = fmaf(v.w - mu.w, sc.w, sh.w);
            }
            if (p.pre_relu) {
              v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
            }
          } else {
            v = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          float4 h, l;
          split_tf32_fast(v.x, h.x, l.x); split_tf32_fast(v.y, h.y, l.y);
          split_tf32_fast(v.z, h.z, l.z); split_tf32_fast(v.w, h.w, l.w);
          sts128(addr, h);
          if (split) sts128(addr + (uint32_t)p.half_bytes, l);
        }
      }
      // ---- dY blocks: split only
      for (int j = 0; j < p.nblk; ++j) {
        const uint32_t addr = base + x_bytes + (uint32_t)(j * kKp * 128) + (uint32_t)r * 128u + (uint32_t)pc * 16u;
        const float4 v = lds128(addr);
        float4 h, l;
        split_tf32_fast(v.x, h.x, l.x); split_tf32_fast(v.y, h.y, l.y);
        split_tf32_fast(v.z, h.z, l.z); split_tf32_fast(v.w, h.w, l.w);
        sts128(addr, h);
        if (split) sts128(addr + (uint32_t)p.half_bytes, l);
      }
      fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&ready_bar[s]);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc_dyn(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// dw[co][ci][tap] = scale * sum_splits partial[split][tap][ci][co]. Block = 32 consecutive partial elements (coalesced
// along co) x 8 split groups, fixed-order final sum (deterministic).
__global__ void __launch_bounds__(256)
wgrad_reduce3_kernel(const float* __restrict__ partial, float* __restrict__ dw, float scale, int Cin, int Cout,
                     int splits) {
  __shared__ float red[8][32];
  const int64_t total = (int64_t)9 * Cin * Cout;
  const int lane = threadIdx.x & 31, sg = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 32 + lane;   // index into [tap][ci][co]
  float acc = 0.f;
  if (i < total) {
#pragma unroll 4
    for (int s = sg; s < splits; s += 8) acc += partial[(size_t)s * total + i];
  }
  red[sg][lane] = acc;
  __syncthreads();
  if (sg != 0 || i >= total) return;
  float t = red[0][lane];
#pragma unroll
  for (int k = 1; k < 8; ++k) t += red[k][lane];
  const int co = (int)(i % Cout);
  const int ci = (int)((i / Cout) % Cin);
  const int tap = (int)(i / ((int64_t)Cout * Cin));
  dw[((size_t)co * Cin + ci) * 9 + tap] = t * scale;
}

struct Plan3 {
  bool ok;
  int bw, bh, halo_w, halo_h, halo_px, cblks, nblk, xblk_bytes, half_bytes, stage_bytes, stages, groups, tg, tmem_cols;
  int pair;
};

Plan3 plan3(int H, int W, int Cin, int Cout, int passes) {
  Plan3 pl{};
  pl.ok = false;
  static const bool pair_ok = [] { const char* e = getenv("FPD_WGRAD3_PAIR"); return !(e && e[0] == '0'); }();
  pl.pair = (Cin == 32 && pair_ok) ? 1 : 0;                    // two taps per M = 64 instruction
  if (!(Cin == 64 || Cin == 128 || pl.pair)) return pl;        // M of the MMA
  if (Cout % 32 != 0 || Cout < 32 || Cout > 256) return pl;    // N of the MMA (multiple of 16), 32-channel TMA boxes
  if (W % 8 != 0) return pl;                                   // a K step = 8 pixels consecutive along w
  int bw = 8;
  while (bw * 2 <= kKp && W % (bw * 2) == 0) bw *= 2;
  const int bh = kKp / bw;
  if (H % bh != 0) return pl;
  pl.bw = bw; pl.bh = bh;
  pl.halo_w = bw + 2; pl.halo_h = bh + 2; pl.halo_px = pl.halo_w * pl.halo_h;
  if (pl.halo_px > 128) return pl;
  pl.cblks = Cin / 32; pl.nblk = Cout / 32;
  pl.xblk_bytes = (pl.halo_px * 128 + 511) / 512 * 512;
  pl.half_bytes = (pl.cblks * pl.xblk_bytes + pl.nblk * kKp * 128 + 1023) / 1024 * 1024;
  pl.stage_bytes = (passes == 3 ? 2 : 1) * pl.half_bytes;
  pl.stages = (226 * 1024 - 1024 - kTail) / pl.stage_bytes;
  if (pl.stages > 8) pl.stages = 8;
  if (pl.stages < 2) return pl;
  pl.groups = (9 * Cout + 511) / 512;
  pl.tg = (9 + pl.groups - 1) / pl.groups;
  if (pl.pair) {                  // five tap pairs, all in one CTA
    pl.groups = 1;
    pl.tg = 5;
  }
  int tc = 32;
  while (tc < pl.tg * Cout) tc *= 2;
  if (tc > 512) return pl;
  pl.tmem_cols = tc;
  pl.ok = true;
  return pl;
}

}  // namespace

bool wgrad_tc3_supported(int H, int W, int Cin, int Cout, int ksize) {
  static const bool off = [] { const char* e = getenv("FPD_WGRAD3"); return e && e[0] == '0'; }();
  return !off && ksize == 3 && plan3(H, W, Cin, Cout, 3).ok;
}

size_t wgrad_tc3_workspace_bytes(int B, int H, int W, int Cin, int Cout, int num_sms) {
  Plan3 pl = plan3(H, W, Cin, Cout, 3);
  if (!pl.ok) return 0;
  const int num_ktiles = (W / pl.bw) * (H / pl.bh) * B;
  int splits = num_sms / pl.groups;
  if (splits < 1) splits = 1;
  if (splits > num_ktiles) splits = num_ktiles;
  return (size_t)splits * 9 * Cin * Cout * sizeof(float);
}

int wgrad_tc3_launch(const float* x, const float* pre_mean, const float* pre_scale, const float* pre_shift,
                     int pre_relu, const float* dy, int passes, float* dw_oihw, float scale, int B, int H, int W,
                     int Cin, int Cout, void* workspace, size_t workspace_bytes, int num_sms, cudaStream_t stream) {
  Plan3 pl = plan3(H, W, Cin, Cout, passes);
  FPD_REQUIRE(pl.ok, "wgrad_tc3: unsupported shape H=%d W=%d Cin=%d Cout=%d", H, W, Cin, Cout);
  FPD_REQUIRE(x && dy && dw_oihw, "wgrad_tc3: null operand");
  FPD_REQUIRE(passes == 1 || passes == 3, "wgrad_tc3: passes must be 1 or 3");
  FPD_REQUIRE((pre_scale == nullptr) == (pre_shift == nullptr), "wgrad_tc3: pre_scale/pre_shift come in pairs");
  Wgrad3Params p{};
  p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.passes = passes;
  p.bw = pl.bw; p.bh = pl.bh; p.tiles_w = W / pl.bw; p.tiles_h = H / pl.bh;
  p.num_ktiles = p.tiles_w * p.tiles_h * B;
  p.halo_w = pl.halo_w; p.halo_h = pl.halo_h; p.halo_px = pl.halo_px;
  p.cblks = pl.cblks; p.nblk = pl.nblk; p.xblk_bytes = pl.xblk_bytes; p.half_bytes = pl.half_bytes;
  p.stage_bytes = pl.stage_bytes; p.stages = pl.stages;
  p.groups = pl.groups; p.tg = pl.tg; p.pair = pl.pair;
  p.splits = num_sms / pl.groups;
  if (p.splits < 1) p.splits = 1;
  if (p.splits > p.num_ktiles) p.splits = p.num_ktiles;
  p.kt_per_split = (p.num_ktiles + p.splits - 1) / p.splits;
  p.tmem_cols = pl.tmem_cols;
  {
    const char* e = getenv("FPD_WGRAD3_LANEMAP");
    p.lane_map = e ? atoi(e) : 1;
  }
  p.pre_mean = pre_mean; p.pre_scale = pre_scale; p.pre_shift = pre_shift; p.pre_relu = pre_relu;
  const size_t need = (size_t)p.splits * 9 * Cin * Cout * sizeof(float);
  FPD_REQUIRE(workspace && workspace_bytes >= need, "wgrad_tc3: workspace too small (%zu < %zu)", workspace_bytes, need);
  p.partial = (float*)workspace;

  CUtensorMap tm_x, tm_g;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t strides[3] = {(uint64_t)Cin * 4, (uint64_t)W * Cin * 4, (uint64_t)H * W * Cin * 4};
    uint32_t box[4] = {32u, (uint32_t)p.halo_w, (uint32_t)p.halo_h, 1u};
    int rc = encode_tmap(&tm_x, x, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t strides[3] = {(uint64_t)Cout * 4, (uint64_t)W * Cout * 4, (uint64_t)H * W * Cout * 4};
    uint32_t box[4] = {32u, (uint32_t)p.bw, (uint32_t)p.bh, 1u};
    int rc = encode_tmap(&tm_g, dy, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    FPD_CUDA_CHECK(cudaFuncSetAttribute(wgrad_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const size_t smem_bytes = (size_t)p.stages * p.stage_bytes + 1024 + kTail;
  wgrad_tc3_kernel<<<p.groups * p.splits, kThreads, smem_bytes, stream>>>(tm_x, tm_g, p);
  FPD_LAUNCH_CHECK();
  const int64_t total = (int64_t)9 * Cin * Cout;
  wgrad_reduce3_kernel<<<(int)((total + 31) / 32), 256, 0, stream>>>(p.partial, dw_oihw, scale, Cin, Cout, p.splits);
  FPD_LAUNCH_CHECK();
  return FPD_OK;
}

}  // namespace fpd
