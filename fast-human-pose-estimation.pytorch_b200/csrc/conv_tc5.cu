// conv_tc5.cu -- implicit-GEMM convolution, generation 5: halo-tile reuse + optional 3xFP16 operands.
//
// What ncu said about conv_tc3.cu (profiles/r1b_prof_conv_tc_ts.md, 3x3 128->128 @64x64, B=32): TMA pulls 1.81 GB
// through the L2->SM crossbar for 67 MB of input (8.7 TB/s, ~3/4 of the measured ~6300 B/clk chip-wide L2 cap):
// every one of the 9 taps re-fetches its shifted x tile (604 MB) and every 128-pixel tile re-fetches the hi/lo fp32
// weight tiles (1.2 GB); and for N <= 64 the eight transform warps, which redo BN-apply + ReLU + split for every tap,
// are the bound (~400 issue cycles per 32-channel stage against 384 cycles of MMA). This kernel changes three things:
//
//  1. 3x3: the x tile is fetched ONCE per channel block with its 1-pixel halo ((bw+2) x (bh+2) x bn pixels, one TMA
//     box, hardware zero fill), transformed + split ONCE into a swizzled shared-memory "split tile", and each of the
//     nine taps is then only a shifted copy shared memory -> tensor memory (8 x ld.shared.v4 + 2 x tcgen05.st per
//     thread and tap, no arithmetic). x traffic drops from 9x to ~1.4x, transform work from 9x to ~1.4x.
//  2. Optional 3xFP16 (mode f16): x = hi + lo with hi, lo in fp16 (11 + 11 significant bits, the same 22 bits 3xTF32
//     keeps), products accumulated in fp32 by tcgen05.mma kind::f16 -- twice the tensor rate of kind::tf32 and half
//     the weight bytes per MAC (weights pre-scaled by 2^8 so their lo parts stay normal; the scale is folded into the
//     epilogue). Channel blocks are 64 wide in this mode (128-byte fp16 rows), 32 in tf32 mode. Used for the forward
//     convolutions; gradients (tiny magnitudes) stay on 3xTF32.
//  3. Separate rings for weights, raw x tiles and TMEM A stages, so the weight stream runs ahead independently.
//
// 1x1 convolutions take the "direct" path (raw tile -> registers -> TMEM, as conv_tc3.cu).
//
// TMEM map (512 columns): [acc 0 | acc 1 | A stages: {hi 32 cols, lo 32 cols} x S_a]. A[m][k] lives at lane m;
// tf32: one value per column; f16: two values per column (even k in the low half).
//
// Warp roles (448 threads, persistent): warp 0 TMA producer, warp 1 TMEM alloc + MMA issuer, warps 2-5 epilogue,
// warps 6-13 transform/copy.
//
// Reference semantics: y = conv(relu?((x - mean) * scale + shift)) (+ bias, + residual), i.e. BatchNorm2d-apply +
// ReLU + Conv2d of lib/models/hourglass.py:34-44 (padding positions are zero AFTER the affine, like Conv2d(padding=1)
// on the activated tensor).
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace fpd {
namespace {

constexpr int kTileM = 128;
constexpr int kThreads = 448;
constexpr int kXf = 256;          // transform threads (warps 6..13)
constexpr int kMaxCinH = 2048;              // pose_resnet's layer4 / first deconv
constexpr int kMinParamFloats = 512 + 64;   // BN parameter arrays in shared memory: sized per launch (ConvHParams::param_floats)
constexpr int kMaxRing = 8;
constexpr int tail_bytes_for(int param_floats) { return 1024 + 3 * param_floats * 4; }
constexpr int kEpiBytes = 4 * 32 * 128;   // epilogue staging: one [32 rows x 128 B] tile per epilogue warp (3x3 path: 4 warps)
// 1x1 ("direct") path: the operand transform is light (raw tile -> registers -> TMEM, no tap copies) while the epilogue
// (TMEM read-out, residual loads, global stores) is what bounds the kernel (r1 stall counters), so the warp budget is
// re-cut there: 4 transform warps (each thread does both 32-channel halves of its row, waiting for the first half's
// tcgen05.st before it reuses the registers) and 8 epilogue warps (warps 2-5 take the even 32-column groups of a tile,
// warps 10-13 the odd ones; a TMEM lane quarter may be read by any warp with the same warp % 4) -- where the epilogue is
// the heavier side; shapes whose input is the wide side keep the 8 + 4 split of the 3x3 path (plan() decides per launch).
constexpr int kEpiBytes8 = 8 * 32 * 128;

struct ConvHParams {
  int B, H, W, Cin, Cout;
  int taps, passes;
  int bn, bh, bw;
  int tiles_w, tiles_h, tiles_n, num_tiles;
  int nt, n_tiles;
  int ncb;                 // channel blocks (64 ch in f16 mode, 32 in tf32 mode)
  int halo_w, halo_h, halo_px;
  int raw_box_bytes;       // bytes of one 32-channel raw box (rounded up to 1024)
  int raw_stage_bytes, raw_stages;
  int raw_halves;          // f16 halo tiles too large for two resident boxes: the two 32-channel boxes of a block pass
                           // through ONE raw buffer one after the other (2), otherwise 1
  int split_bytes;         // 0: direct path
  int w_tile_bytes, w_stage_bytes, w_stages;
  int a_stages, a_col0, tmem_cols, acc_stages;
  const float* pre_mean;
  const float* pre_scale;
  const float* pre_shift;
  int pre_relu;
  const float* bias;
  const float* residual;
  const float* relu_mask;
  float* y;
  float out_scale;
  const float* in_scale;   // nullable device float[2] {S, 1/S}: operands are x * S, the epilogue multiplies by 1/S (exact
                           // powers of two; puts small-magnitude gradients into the fp16 range, see channel_sum amax)
  double* stat_part;        // kStats: per-CTA column sums of the OUTPUT, [gridDim.x][Cout][2] = {sum (y - pivot), sum (y - pivot)^2}
  const float* stat_pivot;  // nullable [Cout]: per-channel pivot (any value near the channel mean; zero if null)
  int stat_bytes;           // shared-memory bytes of the per-warp accumulators (4 or 8 x Cout x 16), 0 without statistics
  int epi8;                 // 1: direct path with 8 epilogue + 4 transform warps
  int epi_bytes;            // staging bytes (kEpiBytes or kEpiBytes8)
  int param_floats;         // length of each of the three BN parameter arrays in shared memory (>= Cin rounded up to kCB)
  long long* prof;   // optional per-CTA stall counters [grid][16] (fpd_conv2d_tc_h_set_profile_buffer); null normally
  int dbg;   // timing ablations only (FPD_CONV_DBG bit mask, tools/diag_conv_h.py): 1 no MMA, 2 no weight TMA, 4 no x TMA,
             // 8 no transform/copy work, 16 no epilogue global traffic, 32 no halo split. Results are garbage when set.
};

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void xf_barrier() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
// One mbarrier arrival per WARP (after every lane has reached this point): 256 per-thread arrivals on one barrier word
// serialise (~0.4 us per pipeline stage measured, tools/diag_conv_h.py --ablate), 8 per-warp arrivals do not.
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

// D[tmem] (+)= A[tmem] * B[smem desc], fp16 inputs, fp32 accumulate.
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Instruction descriptor for kind::f16 (fp16 x fp16 -> fp32): c_format = 1 (F32), a_format = b_format = 0 (F16),
// both operands K-major.
__host__ __device__ constexpr uint32_t umma_idesc_f16(uint32_t M, uint32_t N) {
  return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// (v0, v1) -> packed fp16 hi pair and lo pair: v ~= hi + lo. v0 goes to the low half (even k).
__device__ __forceinline__ void split_f16x2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  v0 = fminf(fmaxf(v0, -65504.f), 65504.f);   // saturate instead of producing inf (never hit by sane activations)
  v1 = fminf(fmaxf(v1, -65504.f), 65504.f);
  const __half2 h = __floats2half2_rn(v0, v1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

__device__ __forceinline__ float4 affine_relu4(float4 v, uint32_t mean_a, uint32_t scale_a, uint32_t shift_a, int c,
                                               bool has_affine, int relu, float in_s) {
  if (has_affine) {
    const float4 mu = lds128(mean_a + (uint32_t)c * 4u);
    const float4 sc = lds128(scale_a + (uint32_t)c * 4u);
    const float4 sh = lds128(shift_a + (uint32_t)c * 4u);
    v.x = fmaf(v.x - mu.x, sc.x, sh.x); v.y = fmaf(v.y - mu.y, sc.y, sh.y);
    v.z = fmaf(v.z - mu.z, sc.z, sh.z); v.w = fmaf(v.w - mu.w, sc.w, sh.w);
  }
  if (relu) {
    v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
  }
  v.x *= in_s; v.y *= in_s; v.z *= in_s; v.w *= in_s;
  return v;
}

// mbarrier wait that accumulates its stall cycles into acc when profiling is on
__device__ __forceinline__ void timed_wait(uint64_t* bar, uint32_t parity, bool prof, long long& acc) {
  if (prof) {
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    acc += clock64() - t0;
  } else {
    mbar_wait(bar, parity);
  }
}

struct TileCoord {
  int w0, h0, n0, n0w;
};
__device__ __forceinline__ TileCoord tile_coord(const ConvHParams& p, int tile) {
  const int mt = tile / p.n_tiles;
  TileCoord t;
  t.w0 = (mt % p.tiles_w) * p.bw;
  t.h0 = ((mt / p.tiles_w) % p.tiles_h) * p.bh;
  t.n0 = (mt / (p.tiles_w * p.tiles_h)) * p.bn;
  t.n0w = (tile % p.n_tiles) * p.nt;
  return t;
}

// kStats: the epilogue also accumulates, per output channel, the sums the next train-mode BatchNorm needs (nn.BatchNorm2d
// batch statistics of this tensor, hourglass.py:18-26) -- the separate statistics pass over the tensor (bn_stats_partial,
// one more HBM read per BatchNorm) disappears. Per lane: fp32 sums over its 8 rows of (y - pivot) and (y - pivot)^2
// (8-term sums: as tight as the pivoted 14-row runs of bn_stats_partial_kernel), then fp64: shuffle tree over the 4 lanes
// sharing a column, per-WARP shared-memory accumulators (no atomics: deterministic), fixed-order merge of the 4 warps
// at the end, one [Cout][2] block of partial sums per CTA; bn_sums_finalize_kernel (elementwise.cu) merges the CTAs.
// A separate instantiation: the plain kernel's code is unchanged.
template <bool kF16, bool kStats>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_h_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w_hi,
                 const __grid_constant__ CUtensorMap tm_w_lo, const ConvHParams p) {
  constexpr int kCB = kF16 ? 64 : 32;      // channels per block
  constexpr int kBoxes = kF16 ? 2 : 1;     // 32-channel fp32 TMA boxes per raw block
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const bool split = p.passes == 3;
  const bool halo = p.split_bytes > 0;
  uint8_t* raw_base = smem;
  uint8_t* split_base = raw_base + (size_t)p.raw_stages * p.raw_stage_bytes;
  uint8_t* w_base = split_base + p.split_bytes;
  uint8_t* epi_base = w_base + (size_t)p.w_stages * p.w_stage_bytes;
  uint8_t* tail = epi_base + p.epi_bytes;
  uint64_t* raw_full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* raw_empty = raw_full + kMaxRing;
  uint64_t* w_full = raw_empty + kMaxRing;
  uint64_t* w_empty = w_full + kMaxRing;
  uint64_t* a_ready = w_empty + kMaxRing;
  uint64_t* a_empty = a_ready + kMaxRing;
  uint64_t* tmem_full = a_empty + kMaxRing;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* s_mean = reinterpret_cast<float*>(tail + 1024);
  float* s_scale = s_mean + p.param_floats;
  float* s_shift = s_scale + p.param_floats;
  double* s_stat = reinterpret_cast<double*>(tail + tail_bytes_for(p.param_floats));   // kStats: [4 epilogue warps][Cout][2]

  // warp index made provably warp-uniform (shfl broadcast): the single-thread TMA / MMA issue code below then keeps its
  // descriptors in uniform registers. With a threadIdx-derived `if (lane == 0)` around the whole role ptxas wraps every
  // UTCHMMA / UTMALDG in an ELECT + R2UR.BROADCAST + BRA.U.ANY waterfall loop (~100 cycles per MMA, measured: the issuing
  // thread, not the tensor core, bounded the kernel).
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w_hi);
    if (split) tma_prefetch_desc(&tm_w_lo);
    const uint32_t n_xf_warps = p.epi8 ? 4u : (uint32_t)(kXf / 32), n_epi_warps = p.epi8 ? 8u : 4u;
    for (int s = 0; s < kMaxRing; ++s) {
      mbar_init(&raw_full[s], 1);
      mbar_init(&raw_empty[s], n_xf_warps);
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
      mbar_init(&a_ready[s], n_xf_warps);
      mbar_init(&a_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], n_epi_warps);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc_dyn(tmem_ptr_smem, 512u);
  const int cin_pad = p.ncb * kCB;
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

  if (warp == 0) {
    // ===================== TMA producer =====================
    // whole warp runs the loop; one elected lane arms the barrier and issues the bulk-tensor copies
    {
      int rs = 0, ws = 0;
      uint32_t rph = 0, wph = 0;
      const uint32_t w_tx = (uint32_t)p.w_stage_bytes;
      const bool prof = p.prof != nullptr;
      long long c_wempty = 0, c_rempty = 0;
      const long long c_start = prof ? clock64() : 0;
      // raw block loader: one (halo) tile of kCB channels as 1-2 boxes of 32 channels
      // hf0..hf1: which fills of the block to issue (sequential mode has two per block; the second one of the NEXT block
      // may only be waited for after every weight tile of the current block is in flight, see below)
      auto issue_raw = [&](int tile, int cb, int hf0, int hf1) {
        const TileCoord t = tile_coord(p, tile);
        const int nbox = (kBoxes == 2 && p.Cin - cb * kCB > 32) ? 2 : 1;
        const uint32_t box_tx = (uint32_t)((halo ? p.halo_px : kTileM) * 128);
        const int wc = halo ? t.w0 - 1 : t.w0, hc = halo ? t.h0 - 1 : t.h0;
        for (int hf = hf0; hf < hf1; ++hf) {
          timed_wait(&raw_empty[rs], rph ^ 1, prof, c_rempty);
          if (elect_one()) {
            uint8_t* dst = raw_base + (size_t)rs * p.raw_stage_bytes;
            if (p.dbg & 4) {
              mbar_arrive(&raw_full[rs]);
            } else if (p.raw_halves == 2) {
              // one box per fill; an absent second box (ragged channel block) is a plain arrival, the transform zeroes it
              if (hf < nbox) {
                mbar_expect_tx(&raw_full[rs], box_tx);
                tma_load_4d(dst, &tm_x, &raw_full[rs], cb * kCB + hf * 32, wc, hc, t.n0);
              } else {
                mbar_arrive(&raw_full[rs]);
              }
            } else {
              mbar_expect_tx(&raw_full[rs], box_tx * (uint32_t)nbox);
              tma_load_4d(dst, &tm_x, &raw_full[rs], cb * kCB, wc, hc, t.n0);
              if (nbox == 2) tma_load_4d(dst + p.raw_box_bytes, &tm_x, &raw_full[rs], cb * kCB + 32, wc, hc, t.n0);
            }
          }
          __syncwarp();
          rs = (rs + 1 == p.raw_stages) ? 0 : rs + 1;
          rph ^= (rs == 0);
        }
      };
      const int pre_tap = min(p.w_stages - 1, p.taps - 1);
      int tile = blockIdx.x, cb = 0;
      bool valid = tile < p.num_tiles;
      if (valid) issue_raw(tile, cb, 0, p.raw_halves);
      while (valid) {
        int ntile = tile, ncb_ = cb + 1;
        if (ncb_ == p.ncb) { ncb_ = 0; ntile += gridDim.x; }
        const bool nvalid = ntile < p.num_tiles;
        const int n0w = (tile % p.n_tiles) * p.nt;
        for (int tap = 0; tap < p.taps; ++tap) {
          timed_wait(&w_empty[ws], wph ^ 1, prof, c_wempty);
          if (elect_one()) {
            uint8_t* st = w_base + (size_t)ws * p.w_stage_bytes;
            if (p.dbg & 2) {
              mbar_arrive(&w_full[ws]);
            } else {
              mbar_expect_tx(&w_full[ws], w_tx);
              tma_load_3d(st, &tm_w_hi, &w_full[ws], cb * kCB, n0w, tap);
              if (split) tma_load_3d(st + p.w_tile_bytes, &tm_w_lo, &w_full[ws], cb * kCB, n0w, tap);
            }
          }
          __syncwarp();
          ws = (ws + 1 == p.w_stages) ? 0 : ws + 1;
          wph ^= (ws == 0);
          if (tap == pre_tap && nvalid) issue_raw(ntile, ncb_, 0, 1);
        }
        // sequential mode: the next block's second fill frees up only after this block's taps have run, i.e. after all
        // of its weight tiles were issued -- waiting for it any earlier would deadlock the weight ring
        if (nvalid && p.raw_halves == 2) issue_raw(ntile, ncb_, 1, 2);
        tile = ntile; cb = ncb_; valid = nvalid;
      }
      if (prof && lane == 0) {
        long long* o = p.prof + (size_t)blockIdx.x * 16;
        o[0] = c_wempty; o[1] = c_rempty; o[2] = clock64() - c_start;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the loop (waits included); one elected lane issues the MMAs and commits.
    {
      const uint32_t idesc = kF16 ? umma_idesc_f16(kTileM, (uint32_t)p.nt) : umma_idesc_tf32(kTileM, (uint32_t)p.nt, 0, 0);
      uint32_t tile_iter = 0;
      int ws = 0, as_ = 0;
      uint32_t wph = 0, aph_ = 0;
      const bool prof = p.prof != nullptr;
      long long c_wfull = 0, c_aready = 0, c_tempty = 0;
      const long long c_start = prof ? clock64() : 0;
      const uint32_t w_base_a = smem_u32(w_base);
      const int ksteps = p.ncb * p.taps;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tile_iter) {
        const uint32_t acs = p.acc_stages == 2 ? (tile_iter & 1) : 0u;
        const uint32_t acph = (p.acc_stages == 2 ? (tile_iter >> 1) : tile_iter) & 1;
        timed_wait(&tmem_empty[acs], acph ^ 1, prof, c_tempty);
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + acs * (uint32_t)p.tmem_cols;
        for (int kb = 0; kb < ksteps; ++kb) {
          timed_wait(&w_full[ws], wph, prof, c_wfull);
          timed_wait(&a_ready[as_], aph_, prof, c_aready);
          tc_fence_after_sync();
          if (elect_one()) {
            const uint32_t b_hi = w_base_a + (uint32_t)ws * (uint32_t)p.w_stage_bytes;
            const uint64_t db_hi0 = umma_desc_sw128(b_hi, 16, 1024);
            const uint64_t db_lo0 = umma_desc_sw128(b_hi + (uint32_t)p.w_tile_bytes, 16, 1024);
            const uint32_t ta_hi = tmem_base + (uint32_t)p.a_col0 + (uint32_t)as_ * 64u;
            const uint32_t ta_lo = ta_hi + 32u;
            if (!(p.dbg & 1)) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                // +32 bytes along K inside the 128-byte swizzle row = +2 in the (address >> 4) start field
                const uint64_t db_hi = db_hi0 + (uint64_t)(ks * 2);
                uint32_t acc = (kb > 0 || ks > 0) ? 1u : 0u;
                if (split) {
                  const uint64_t db_lo = db_lo0 + (uint64_t)(ks * 2);
                  if (kF16) {
                    umma_f16_ts(tmem_d, ta_lo + ks * 8, db_hi, idesc, acc);
                    umma_f16_ts(tmem_d, ta_hi + ks * 8, db_lo, idesc, 1u);
                  } else {
                    umma_tf32_ts(tmem_d, ta_lo + ks * 8, db_hi, idesc, acc);
                    umma_tf32_ts(tmem_d, ta_hi + ks * 8, db_lo, idesc, 1u);
                  }
                  acc = 1u;
                }
                if (kF16) umma_f16_ts(tmem_d, ta_hi + ks * 8, db_hi, idesc, acc);
                else umma_tf32_ts(tmem_d, ta_hi + ks * 8, db_hi, idesc, acc);
              }
            }
            umma_commit(&w_empty[ws]);
            umma_commit(&a_empty[as_]);
          }
          __syncwarp();
          ws = (ws + 1 == p.w_stages) ? 0 : ws + 1;
          wph ^= (ws == 0);
          as_ = (as_ + 1 == p.a_stages) ? 0 : as_ + 1;
          aph_ ^= (as_ == 0);
        }
        if (elect_one()) umma_commit(&tmem_full[acs]);
        __syncwarp();
      }
      if (prof && lane == 0) {
        long long* o = p.prof + (size_t)blockIdx.x * 16;
        o[3] = c_wfull; o[4] = c_aready; o[5] = c_tempty; o[6] = clock64() - c_start;
      }
    }
  } else if (warp < 6 || (p.epi8 && warp >= 10)) {
    // ===================== epilogue =====================
    // TMEM -> registers (thread = pixel row) -> per-warp shared-memory staging tile [32 rows x 128 B, chunk-swizzled]
    // -> coalesced global traffic (8 lanes cover one 128-byte row segment: 4 full lines per instruction instead of 32
    // partial ones; the per-thread-row form cost 8k L1 wavefronts per tile and was THE bound of the 1x1 convolutions,
    // tools/diag_conv_h.py --ablate). Residual / bias / mask are applied in the coalesced phase.
    const int q = warp & 3;                                   // TMEM lane quarter
    const int e = warp < 6 ? warp - 2 : warp - 10 + 4;        // epilogue warp index 0..7 (4..7 only with epi8)
    const int cg0 = p.epi8 ? (e >> 2) * 32 : 0;               // first 32-column group of a tile this warp takes
    const int cgs = p.epi8 ? 64 : 32;                         // ... and its stride
    const uint32_t stg = smem_u32(epi_base) + (uint32_t)e * 4096u;
    const int sub = lane >> 3, ch = lane & 7;
    const float oscale = p.in_scale ? p.out_scale * __ldg(p.in_scale + 1) : p.out_scale;
    uint32_t tile_iter = 0;
    const bool prof = p.prof != nullptr;
    long long c_tfull = 0;
    const long long c_start = prof ? clock64() : 0;
    double* my_stat = s_stat + (size_t)e * p.Cout * 2;   // this warp's accumulators
    if (kStats) {
      for (int i = lane; i < p.Cout * 2; i += 32) my_stat[i] = 0.0;
      __syncwarp();
    }
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tile_iter) {
      const uint32_t acs = p.acc_stages == 2 ? (tile_iter & 1) : 0u;
      const uint32_t acph = (p.acc_stages == 2 ? (tile_iter >> 1) : tile_iter) & 1;
      const TileCoord t = tile_coord(p, tile);
      // rows this lane writes in the coalesced phase: m = 32 q + 4 it + sub
      uint32_t pixoff[8];   // element offsets (the launcher checks that the output has < 2^32 elements)
      uint32_t vmask = 0;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int m = q * 32 + it * 4 + sub;
        const int pw = t.w0 + (m % p.bw);
        const int ph_ = t.h0 + (m / p.bw) % p.bh;
        const int pn = t.n0 + m / (p.bw * p.bh);
        if (pn < p.B) vmask |= 1u << it;
        pixoff[it] = (uint32_t)((((size_t)pn * p.H + ph_) * p.W + pw) * (size_t)p.Cout + (size_t)t.n0w);
      }
      if (p.dbg & 16) vmask = 0;
      // Pull the residual rows towards the L2 one tile ahead (this tile's too, the first time): the loads below then
      // cost an L2 hit instead of an HBM round trip per 32-column group, which was what paced the 1x1 convolutions
      // (epilogue busy 6 us per tile against 3 us of HBM time, tools/diag_conv_h.py --stalls).
      if (p.residual && !(p.dbg & 16) && ch == 0 && e < 4) {
        for (int which = (tile_iter == 0 ? 0 : 1); which < 2; ++which) {
          const int tl = tile + which * (int)gridDim.x;
          if (tl >= p.num_tiles) break;
          const TileCoord tp = tile_coord(p, tl);
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int m = q * 32 + it * 4 + sub;
            const int pn = tp.n0 + m / (p.bw * p.bh);
            if (pn >= p.B) continue;
            const size_t off = (((size_t)pn * p.H + tp.h0 + (m / p.bw) % p.bh) * p.W + tp.w0 + (m % p.bw)) * (size_t)p.Cout +
                               (size_t)tp.n0w;
            for (int c = 0; c < p.nt; c += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.residual + off + c));
          }
        }
      }
      timed_wait(&tmem_full[acs], acph, prof, c_tfull);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + acs * (uint32_t)p.tmem_cols + ((uint32_t)(q * 32) << 16);
      for (int c0 = cg0; c0 < p.nt; c0 += cgs) {
        const int gw = min(32, p.nt - c0);   // 32, or 16 for the last group of a slice that is not a multiple of 32
        const int col = c0 + ch * 4;
        const bool act = ch * 4 < gw;
        // residual rows of this group: issued before the TMEM read-out so their latency overlaps it (and kept apart from
        // the stores below -- y and residual may alias as far as the compiler knows, which would serialise load/store pairs)
        float4 r4[8];
        if (p.residual && act) {
#pragma unroll
          for (int it = 0; it < 8; ++it)
            if (vmask & (1u << it)) r4[it] = __ldg(reinterpret_cast<const float4*>(p.residual + (size_t)pixoff[it] + col));
        }
        {
          uint32_t v[16];
          tmem_ld16(taddr + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float4 o;
            o.x = __uint_as_float(v[4 * j + 0]) * oscale; o.y = __uint_as_float(v[4 * j + 1]) * oscale;
            o.z = __uint_as_float(v[4 * j + 2]) * oscale; o.w = __uint_as_float(v[4 * j + 3]) * oscale;
            sts128(stg + (uint32_t)lane * 128u + (uint32_t)((j ^ (lane & 7)) << 4), o);
          }
          if (gw > 16) {
            tmem_ld16(taddr + c0 + 16, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float4 o;
              o.x = __uint_as_float(v[4 * j + 0]) * oscale; o.y = __uint_as_float(v[4 * j + 1]) * oscale;
              o.z = __uint_as_float(v[4 * j + 2]) * oscale; o.w = __uint_as_float(v[4 * j + 3]) * oscale;
              sts128(stg + (uint32_t)lane * 128u + (uint32_t)(((j + 4) ^ (lane & 7)) << 4), o);
            }
          }
        }
        __syncwarp();
        float ps1[4] = {0.f, 0.f, 0.f, 0.f}, ps2[4] = {0.f, 0.f, 0.f, 0.f};   // kStats: this lane's 8-row sums
        if (act) {
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + t.n0w + col));
          float4 pv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (kStats && p.stat_pivot) pv = __ldg(reinterpret_cast<const float4*>(p.stat_pivot + t.n0w + col));
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rl = it * 4 + sub;
            float4 o = lds128(stg + (uint32_t)rl * 128u + (uint32_t)((ch ^ (rl & 7)) << 4));
            if (vmask & (1u << it)) {
              const size_t off = (size_t)pixoff[it] + (size_t)col;
              o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
              if (p.residual) {
                o.x += r4[it].x; o.y += r4[it].y; o.z += r4[it].z; o.w += r4[it].w;
              }
              if (p.relu_mask) {
                const float4 k = __ldg(reinterpret_cast<const float4*>(p.relu_mask + off));
                o.x = k.x > 0.f ? o.x : 0.f; o.y = k.y > 0.f ? o.y : 0.f;
                o.z = k.z > 0.f ? o.z : 0.f; o.w = k.w > 0.f ? o.w : 0.f;
              }
              *reinterpret_cast<float4*>(p.y + off) = o;
              if (kStats) {
                const float d0 = o.x - pv.x, d1 = o.y - pv.y, d2 = o.z - pv.z, d3 = o.w - pv.w;
                ps1[0] += d0; ps1[1] += d1; ps1[2] += d2; ps1[3] += d3;
                ps2[0] = fmaf(d0, d0, ps2[0]); ps2[1] = fmaf(d1, d1, ps2[1]);
                ps2[2] = fmaf(d2, d2, ps2[2]); ps2[3] = fmaf(d3, d3, ps2[3]);
              }
            }
          }
        }
        if (kStats) {
          // fp64 from here on: sum over the 4 lanes (sub = 0..3) that hold the same 4 columns, then into this warp's
          // accumulators (lanes 0..7 own 4 distinct columns each: no conflicts, no atomics)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            double a = (double)ps1[j], b = (double)ps2[j];
            a += __shfl_xor_sync(0xffffffffu, a, 8);  b += __shfl_xor_sync(0xffffffffu, b, 8);
            a += __shfl_xor_sync(0xffffffffu, a, 16); b += __shfl_xor_sync(0xffffffffu, b, 16);
            if (sub == 0 && act) {
              double* dst = my_stat + (size_t)(t.n0w + col + j) * 2;
              dst[0] += a;
              dst[1] += b;
            }
          }
        }
        __syncwarp();   // staging tile is rewritten by the next column group
      }
      tc_fence_before_sync();
      warp_arrive(&tmem_empty[acs], lane);
    }
    if (kStats) {
      // merge the four warps in a fixed order and publish this CTA's partial sums
      const int nepi = p.epi8 ? 8 : 4;
      if (p.epi8) asm volatile("bar.sync 2, 256;" ::: "memory");
      else asm volatile("bar.sync 2, 128;" ::: "memory");
      const int et = e * 32 + lane;
      for (int i = et; i < p.Cout * 2; i += nepi * 32) {
        double v = 0.0;
        for (int w8 = 0; w8 < nepi; ++w8) v += s_stat[(size_t)w8 * p.Cout * 2 + i];   // fixed order
        p.stat_part[(size_t)blockIdx.x * p.Cout * 2 + i] = v;
      }
    }
    if (prof && warp == 2 && lane == 0) {
      long long* o = p.prof + (size_t)blockIdx.x * 16;
      o[7] = c_tfull; o[8] = clock64() - c_start;
    }
  } else {
    // ===================== operand transform / tap copy -> TMEM =====================
    const int xt = threadIdx.x - 6 * 32;         // 0..255
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int half = (warp - 6) >> 2;             // which half of the block's channels (columns) this thread writes
    const int r = q * 32 + lane;                  // output pixel row of the tile = TMEM lane
    const int dn = r / (p.bw * p.bh), dh_ = (r / p.bw) % p.bh, dw_ = r % p.bw;
    const bool has_affine = p.pre_scale != nullptr;
    const float in_s = p.in_scale ? __ldg(p.in_scale) : 1.f;
    const uint32_t mean_a = smem_u32(s_mean), scale_a = smem_u32(s_scale), shift_a = smem_u32(s_shift);
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)p.a_col0 + (uint32_t)half * 16u;
    int rs = 0, as_ = 0;
    uint32_t rph = 0, aph_ = 0;
    const bool prof = p.prof != nullptr;
    long long c_rfull = 0, c_aempty = 0, c_bar = 0;
    const long long c_start = prof ? clock64() : 0;

    if (!halo) {
      // ---------- direct path (1x1): raw tile -> registers -> TMEM ----------
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const TileCoord t = tile_coord(p, tile);
        const bool inb = t.n0 + dn < p.B;
        for (int cb = 0; cb < p.ncb; ++cb) {
          timed_wait(&raw_full[rs], rph, prof, c_rfull);
          timed_wait(&a_empty[as_], aph_ ^ 1, prof, c_aempty);
          tc_fence_after_sync();
          const uint32_t rawst = smem_u32(raw_base + (size_t)rs * p.raw_stage_bytes);
          // epi8: this thread prepares both 32-channel halves of its row, one after the other
          const int h_lo = p.epi8 ? 0 : half, h_hi = p.epi8 ? 2 : half + 1;
          for (int hh = h_lo; hh < h_hi; ++hh) {
          uint32_t hi[16], lo[16];
          if (kF16) {
            // box `hh` (channels 32*hh .. +31 of the block), all 8 chunks of this thread's 128-byte row
            const bool box_ok = inb && (hh == 0 || p.Cin - cb * kCB > 32) && !(p.dbg & 8);
            const uint32_t xrow = rawst + (uint32_t)hh * (uint32_t)p.raw_box_bytes + (uint32_t)r * 128u;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
              if (box_ok) {
                v = lds128(xrow + (uint32_t)((i ^ (r & 7)) << 4));
                v = affine_relu4(v, mean_a, scale_a, shift_a, cb * kCB + hh * 32 + i * 4, has_affine, p.pre_relu, in_s);
              }
              split_f16x2(v.x, v.y, hi[2 * i], lo[2 * i]);
              split_f16x2(v.z, v.w, hi[2 * i + 1], lo[2 * i + 1]);
            }
          } else {
            const uint32_t xrow = rawst + (uint32_t)r * 128u;
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
              const int i = hh * 4 + ii;
              float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
              if (inb && !(p.dbg & 8)) {
                v = lds128(xrow + (uint32_t)((i ^ (r & 7)) << 4));
                v = affine_relu4(v, mean_a, scale_a, shift_a, cb * kCB + i * 4, has_affine, p.pre_relu, in_s);
              }
              float4 h, l;
              split_tf32_fast(v.x, h.x, l.x); split_tf32_fast(v.y, h.y, l.y);
              split_tf32_fast(v.z, h.z, l.z); split_tf32_fast(v.w, h.w, l.w);
              hi[ii * 4 + 0] = __float_as_uint(h.x); hi[ii * 4 + 1] = __float_as_uint(h.y);
              hi[ii * 4 + 2] = __float_as_uint(h.z); hi[ii * 4 + 3] = __float_as_uint(h.w);
              lo[ii * 4 + 0] = __float_as_uint(l.x); lo[ii * 4 + 1] = __float_as_uint(l.y);
              lo[ii * 4 + 2] = __float_as_uint(l.z); lo[ii * 4 + 3] = __float_as_uint(l.w);
            }
          }
          const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)p.a_col0 + (uint32_t)hh * 16u +
                              (uint32_t)as_ * 64u;
          if (!(p.dbg & 8)) {
            tmem_st16(ta, hi);
            if (split) tmem_st16(ta + 32u, lo);
            // tcgen05.st is asynchronous: its source registers must stay untouched until tcgen05.wait::st -- the second
            // half reuses them, so wait here (without this the A tiles were corrupted whenever the store pipe lagged,
            // i.e. only inside the concurrent multi-stream step, never kernel by kernel)
            if (hh + 1 < h_hi) tmem_st_wait();
          }
          }   // halves
          warp_arrive(&raw_empty[rs], lane);    // every lane's reads of the raw stage have been consumed by the stores above
          if (!(p.dbg & 8)) tmem_st_wait();
          tc_fence_before_sync();
          warp_arrive(&a_ready[as_], lane);
          rs = (rs + 1 == p.raw_stages) ? 0 : rs + 1;
          rph ^= (rs == 0);
          as_ = (as_ + 1 == p.a_stages) ? 0 : as_ + 1;
          aph_ ^= (as_ == 0);
        }
      }
    } else {
      // ---------- halo path (3x3): split once per block, then nine shifted copies ----------
      const uint32_t split_a = smem_u32(split_base);
      const int pc = (dn * p.halo_h + dh_ + 1) * p.halo_w + dw_ + 1;   // halo index of this thread's centre pixel
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const TileCoord t = tile_coord(p, tile);
        // validity of the (up to two) halo pixels this thread splits: inside the image and the batch
        bool ok0 = false, ok1 = false;
        {
          const int hp = p.halo_w * p.halo_h;
          int pp = xt;
          if (pp < p.halo_px) {
            const int n_ = pp / hp, rem = pp - n_ * hp, y_ = rem / p.halo_w, x_ = rem - y_ * p.halo_w;
            const int hh = t.h0 - 1 + y_, ww = t.w0 - 1 + x_;
            ok0 = (t.n0 + n_ < p.B) && hh >= 0 && hh < p.H && ww >= 0 && ww < p.W;
          }
          pp = xt + kXf;
          if (pp < p.halo_px) {
            const int n_ = pp / hp, rem = pp - n_ * hp, y_ = rem / p.halo_w, x_ = rem - y_ * p.halo_w;
            const int hh = t.h0 - 1 + y_, ww = t.w0 - 1 + x_;
            ok1 = (t.n0 + n_ < p.B) && hh >= 0 && hh < p.H && ww >= 0 && ww < p.W;
          }
        }
        for (int cb = 0; cb < p.ncb; ++cb) {
          const bool box1 = kBoxes == 2 && p.Cin - cb * kCB > 32;
          for (int hf = 0; hf < p.raw_halves; ++hf) {
          timed_wait(&raw_full[rs], rph, prof, c_rfull);
          if (hf == 0) {
            const long long t0 = prof ? clock64() : 0;
            xf_barrier();   // every thread has finished the tap copies of the previous block: the split tile is free
            if (prof) c_bar += clock64() - t0;
          }
          const uint32_t rawst = smem_u32(raw_base + (size_t)rs * p.raw_stage_bytes);
          const int g_lo = p.raw_halves == 2 ? hf * 4 : 0, g_hi = p.raw_halves == 2 ? hf * 4 + 4 : 8;
#pragma unroll 1
          for (int g = g_lo; g < g_hi; ++g) {
            if (p.dbg & 32) break;
#pragma unroll
            for (int it = 0; it < 2; ++it) {
              const int pp = xt + it * kXf;
              if (pp >= p.halo_px) break;
              const bool ok = it == 0 ? ok0 : ok1;
              const uint32_t sw = (uint32_t)(pp & 7);
              uint4 oh = make_uint4(0u, 0u, 0u, 0u), ol = make_uint4(0u, 0u, 0u, 0u);
              if (kF16) {
                const int b = g >> 2;
                if (ok && (b == 0 || box1)) {
                  // sequential mode: the box of this half always sits at the start of the (single) raw buffer
                  const uint32_t xrow = rawst + (p.raw_halves == 2 ? 0u : (uint32_t)b * (uint32_t)p.raw_box_bytes) +
                                        (uint32_t)pp * 128u;
                  const uint32_t c0 = (uint32_t)((2 * g) & 7);
                  float4 v0 = lds128(xrow + ((c0 ^ sw) << 4));
                  float4 v1 = lds128(xrow + (((c0 + 1) ^ sw) << 4));
                  const int c = cb * kCB + g * 8;
                  v0 = affine_relu4(v0, mean_a, scale_a, shift_a, c, has_affine, p.pre_relu, in_s);
                  v1 = affine_relu4(v1, mean_a, scale_a, shift_a, c + 4, has_affine, p.pre_relu, in_s);
                  split_f16x2(v0.x, v0.y, oh.x, ol.x);
                  split_f16x2(v0.z, v0.w, oh.y, ol.y);
                  split_f16x2(v1.x, v1.y, oh.z, ol.z);
                  split_f16x2(v1.z, v1.w, oh.w, ol.w);
                }
              } else {
                if (ok) {
                  const uint32_t xrow = rawst + (uint32_t)pp * 128u;
                  float4 v = lds128(xrow + (((uint32_t)g ^ sw) << 4));
                  v = affine_relu4(v, mean_a, scale_a, shift_a, cb * kCB + g * 4, has_affine, p.pre_relu, in_s);
                  float4 h, l;
                  split_tf32_fast(v.x, h.x, l.x); split_tf32_fast(v.y, h.y, l.y);
                  split_tf32_fast(v.z, h.z, l.z); split_tf32_fast(v.w, h.w, l.w);
                  oh = make_uint4(__float_as_uint(h.x), __float_as_uint(h.y), __float_as_uint(h.z), __float_as_uint(h.w));
                  ol = make_uint4(__float_as_uint(l.x), __float_as_uint(l.y), __float_as_uint(l.z), __float_as_uint(l.w));
                }
              }
              const uint32_t dst = split_a + (uint32_t)pp * 256u + ((((uint32_t)g) ^ sw) << 4);
              sts128u(dst, oh);
              if (split) sts128u(dst + 128u, ol);
            }
          }
          warp_arrive(&raw_empty[rs], lane);   // this warp no longer reads the raw stage
          rs = (rs + 1 == p.raw_stages) ? 0 : rs + 1;
          rph ^= (rs == 0);
          }   // halves
          xf_barrier();   // split tile complete and visible to all copy threads
          for (int tap = 0; tap < 9; ++tap) {
            const int pt = pc + (tap / 3 - 1) * p.halo_w + (tap % 3 - 1);
            const uint32_t srow = split_a + (uint32_t)pt * 256u;
            const uint32_t sw = (uint32_t)(pt & 7);
            timed_wait(&a_empty[as_], aph_ ^ 1, prof, c_aempty);
            tc_fence_after_sync();
            uint32_t hi[16], lo[16];
            if (p.dbg & 8) {
              tc_fence_before_sync();
              warp_arrive(&a_ready[as_], lane);
              as_ = (as_ + 1 == p.a_stages) ? 0 : as_ + 1;
              aph_ ^= (as_ == 0);
              continue;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t off = (((uint32_t)(half * 4 + j)) ^ sw) << 4;
              const uint4 a = lds128u(srow + off);
              hi[4 * j + 0] = a.x; hi[4 * j + 1] = a.y; hi[4 * j + 2] = a.z; hi[4 * j + 3] = a.w;
              if (split) {
                const uint4 b = lds128u(srow + 128u + off);
                lo[4 * j + 0] = b.x; lo[4 * j + 1] = b.y; lo[4 * j + 2] = b.z; lo[4 * j + 3] = b.w;
              }
            }
            const uint32_t ta = lane_base + (uint32_t)as_ * 64u;
            tmem_st16(ta, hi);
            if (split) tmem_st16(ta + 32u, lo);
            tmem_st_wait();
            tc_fence_before_sync();
            warp_arrive(&a_ready[as_], lane);
            as_ = (as_ + 1 == p.a_stages) ? 0 : as_ + 1;
            aph_ ^= (as_ == 0);
          }
        }
      }
    }
    if (prof && warp == 6 && lane == 0) {
      long long* o = p.prof + (size_t)blockIdx.x * 16;
      o[9] = c_rfull; o[10] = c_aempty; o[11] = c_bar; o[12] = clock64() - c_start;
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc_dyn(tmem_base, 512u);
  }
}

long long* g_prof_buf = nullptr;

int pow2_floor_div(int x, int cap) {
  int r = 1;
  while (r * 2 <= cap && x % (r * 2) == 0) r *= 2;
  return r;
}

int align1024(int x) { return (x + 1023) / 1024 * 1024; }

// Fills the geometry / ring sizes; returns false if the shape does not fit.
// epi_inputs: bit 0 = the epilogue adds a residual, bit 1 = it applies a ReLU mask (each one more tensor it has to load).
bool plan(ConvHParams& p, int B, int H, int W, int Cin, int Cout, int ksize, int f16, int passes, int stats = 0,
          int epi_inputs = 0) {
  // Warp split of the 1x1 path, per shape (profiles/r2_conv1x1_split.txt, both splits timed on every 1x1 shape of the two
  // hourglasses): the transform's work per tile grows with Cin, the epilogue's with Cout and with every extra tensor it
  // reads. 4 transform + 8 epilogue warps win where (1 + 1.5 residual + mask) x Cout > Cin (128->256 + residual: 108 vs
  // 135 us; 64->128: 31 vs 38), 8 + 4 where the input is the wide side (256->128: 56 vs 70 us; 256->256: 95 vs 123;
  // 256->16: 44 vs 60). FPD_CONV_EPI8 = 0 / 1 forces one split everywhere.
  static const int epi8_mode = [] { const char* e = getenv("FPD_CONV_EPI8"); return e ? (e[0] == '0' ? 0 : 1) : -1; }();
  const bool epi8_pays = Cout * (2 + 3 * (epi_inputs & 1) + 2 * ((epi_inputs >> 1) & 1)) > 2 * Cin;
  p.epi8 = (ksize == 1 && (epi8_mode == 1 || (epi8_mode < 0 && epi8_pays))) ? 1 : 0;
  p.epi_bytes = p.epi8 ? kEpiBytes8 : kEpiBytes;
  p.stat_bytes = 0;
  if (stats) {
    if (Cout > 256) return false;
    p.stat_bytes = (p.epi8 ? 8 : 4) * Cout * 16;
  }
  const int cbch = f16 ? 64 : 32;
  if (!(ksize == 1 || ksize == 3)) return false;
  if (Cin < 4 || Cin > kMaxCinH || Cin % (f16 ? 8 : 4) != 0) return false;
  if (Cout % 16 != 0 || Cout < 16 || Cout > 2048) return false;
  p.param_floats = Cin <= 512 ? kMinParamFloats : (Cin + 63) / 64 * 64 + 64;
  const int kTailBytes = tail_bytes_for(p.param_floats);
  const int nt = conv_tc_ts_slice(Cout);
  if (nt <= 0) return false;
  p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.taps = ksize * ksize;
  p.passes = passes;
  p.nt = nt;
  p.n_tiles = Cout / nt;
  p.ncb = (Cin + cbch - 1) / cbch;
  if (ksize == 3) {
    p.bw = pow2_floor_div(W, 16);
    p.bh = pow2_floor_div(H, kTileM / p.bw);
  } else {
    p.bw = pow2_floor_div(W, kTileM);
    p.bh = pow2_floor_div(H, kTileM / p.bw);
  }
  p.bn = kTileM / (p.bw * p.bh);
  if (p.bn > 256) return false;
  p.tiles_w = W / p.bw; p.tiles_h = H / p.bh; p.tiles_n = (B + p.bn - 1) / p.bn;
  p.num_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles;
  int tc = 32;
  while (tc < nt) tc *= 2;
  p.tmem_cols = tc;
  p.acc_stages = (2 * tc + 2 * 64 <= 512) ? 2 : 1;
  p.a_col0 = p.acc_stages * tc;
  p.a_stages = (512 - p.a_col0) / 64;
  if (p.a_stages > 6) p.a_stages = 6;
  if (p.a_stages < 2) return false;
  p.w_tile_bytes = nt * 128;
  p.w_stage_bytes = (passes == 3 ? 2 : 1) * p.w_tile_bytes;
  const int budget = 227 * 1024 - 1024 - kTailBytes - p.epi_bytes - p.stat_bytes;
  const int boxes = f16 ? 2 : 1;
  if (ksize == 3) {
    p.halo_w = p.bw + 2; p.halo_h = p.bh + 2;
    p.halo_px = p.halo_w * p.halo_h * p.bn;
    if (p.halo_px > 2 * kXf || p.halo_w > 256 || p.halo_h > 256) return false;
    p.raw_box_bytes = align1024(p.halo_px * 128);
    p.raw_stage_bytes = boxes * p.raw_box_bytes;
    p.split_bytes = align1024(p.halo_px * 256);
    p.raw_stages = 1;
    p.raw_halves = 1;
    int rest = budget - p.split_bytes - p.raw_stage_bytes;
    if (boxes == 2 && rest < 3 * p.w_stage_bytes) {
      // large halo tiles (4x4 / 8x6 images: most of the tile is padding): stream the two 32-channel boxes of a block
      // through one raw buffer instead of keeping both resident
      p.raw_halves = 2;
      p.raw_stage_bytes = p.raw_box_bytes;
      rest = budget - p.split_bytes - p.raw_stage_bytes;
    }
    if (rest < 2 * p.w_stage_bytes) return false;
    p.w_stages = rest / p.w_stage_bytes;
    if (p.w_stages > 4) p.w_stages = 4;
    rest -= p.w_stages * p.w_stage_bytes;
    if (p.w_stages >= 3 && rest >= p.raw_stage_bytes && p.raw_halves == 1) p.raw_stages = 2;
  } else {
    p.halo_w = p.halo_h = p.halo_px = 0;
    p.split_bytes = 0;
    p.raw_halves = 1;
    p.raw_box_bytes = kTileM * 128;
    p.raw_stage_bytes = boxes * p.raw_box_bytes;
    p.w_stages = 3;
    int rest = budget - p.w_stages * p.w_stage_bytes;
    p.raw_stages = rest / p.raw_stage_bytes;
    if (p.raw_stages > 4) p.raw_stages = 4;
    if (p.raw_stages < 2) return false;
    rest -= p.raw_stages * p.raw_stage_bytes;
    if (rest >= p.w_stage_bytes) p.w_stages = 4;
  }
  return p.w_stages <= kMaxRing && p.raw_stages <= kMaxRing && p.a_stages <= kMaxRing;
}

}  // namespace

void conv_tc_h_set_profile_buffer(long long* buf) { g_prof_buf = buf; }

bool conv_tc_h_supported(int Cin, int Cout, int ksize, int H, int W, int f16) {
  ConvHParams p{};
  return plan(p, 1, H, W, Cin, Cout, ksize, f16, 3);
}

int conv_tc_h_stats_grid(int B, int H, int W, int Cin, int Cout, int ksize, int f16, int num_sms) {
  ConvHParams p{};
  if (!plan(p, B, H, W, Cin, Cout, ksize, f16, 3, 1)) return 0;
  return p.num_tiles < num_sms ? p.num_tiles : num_sms;
}

int conv_tc_h_launch(const float* x, const float* pre_mean, const float* pre_scale, const float* pre_shift,
                     int pre_relu, const void* w_hi, const void* w_lo, int f16, const float* bias,
                     const float* residual, const float* relu_mask, float* y, float out_scale, const float* in_scale,
                     int B, int H, int W, int Cin, int Cout, int ksize, int num_sms, cudaStream_t stream,
                     double* stat_part, const float* stat_pivot) {
  FPD_REQUIRE(x && w_hi && y, "conv_tc_h: null operand");
  FPD_REQUIRE((double)B * H * W * Cout < 4294967296.0, "conv_tc_h: output has 2^32 or more elements");
  FPD_REQUIRE((pre_scale == nullptr) == (pre_shift == nullptr), "conv_tc_h: pre_scale/pre_shift come in pairs");
  FPD_REQUIRE(pre_scale != nullptr || pre_mean == nullptr, "conv_tc_h: pre_mean needs pre_scale/pre_shift");
  ConvHParams p{};
  FPD_REQUIRE(plan(p, B, H, W, Cin, Cout, ksize, f16, w_lo ? 3 : 1, stat_part ? 1 : 0,
                   (residual ? 1 : 0) | (relu_mask ? 2 : 0)),
              "conv_tc_h: unsupported shape Cin=%d Cout=%d k=%d H=%d W=%d f16=%d stats=%d", Cin, Cout, ksize, H, W, f16,
              stat_part ? 1 : 0);
  p.stat_part = stat_part;
  p.stat_pivot = stat_pivot;
  p.pre_mean = pre_mean; p.pre_scale = pre_scale; p.pre_shift = pre_shift; p.pre_relu = pre_relu;
  p.bias = bias; p.residual = residual; p.relu_mask = relu_mask; p.y = y;
  p.out_scale = f16 ? out_scale * (1.0f / (float)(1 << kF16WeightScaleLog2)) : out_scale;
  {
    const char* e = getenv("FPD_CONV_DBG");
    p.dbg = e ? atoi(e) : 0;
  }
  p.prof = g_prof_buf;
  p.in_scale = in_scale;
  const size_t smem_bytes = (size_t)p.raw_stages * p.raw_stage_bytes + p.split_bytes +
                            (size_t)p.w_stages * p.w_stage_bytes + p.epi_bytes + tail_bytes_for(p.param_floats) +
                            p.stat_bytes + 1024;
  FPD_REQUIRE(smem_bytes <= 227 * 1024, "conv_tc_h: shared memory plan %zu B too large", smem_bytes);

  CUtensorMap tm_x, tm_w_hi, tm_w_lo;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t strides[3] = {(uint64_t)Cin * 4, (uint64_t)W * Cin * 4, (uint64_t)H * W * Cin * 4};
    uint32_t box[4] = {32u, (uint32_t)(ksize == 3 ? p.halo_w : p.bw), (uint32_t)(ksize == 3 ? p.halo_h : p.bh),
                       (uint32_t)p.bn};
    int rc = encode_tmap(&tm_x, x, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    const uint64_t es = f16 ? 2 : 4;
    uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, (uint64_t)p.taps};
    uint64_t strides[2] = {(uint64_t)Cin * es, (uint64_t)Cout * Cin * es};
    uint32_t box[3] = {(uint32_t)(f16 ? 64 : 32), (uint32_t)p.nt, 1};
    int rc = encode_tmap_dt(&tm_w_hi, w_hi, f16 ? 1 : 0, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = encode_tmap_dt(&tm_w_lo, w_lo ? w_lo : w_hi, f16 ? 1 : 0, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    FPD_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_h_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    FPD_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_h_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    FPD_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_h_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    FPD_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_h_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  if (stat_part) {
    if (f16) conv_tc_h_kernel<true, true><<<grid, kThreads, smem_bytes, stream>>>(tm_x, tm_w_hi, tm_w_lo, p);
    else conv_tc_h_kernel<false, true><<<grid, kThreads, smem_bytes, stream>>>(tm_x, tm_w_hi, tm_w_lo, p);
  } else if (f16) conv_tc_h_kernel<true, false><<<grid, kThreads, smem_bytes, stream>>>(tm_x, tm_w_hi, tm_w_lo, p);
  else conv_tc_h_kernel<false, false><<<grid, kThreads, smem_bytes, stream>>>(tm_x, tm_w_hi, tm_w_lo, p);
  FPD_LAUNCH_CHECK();
  return FPD_OK;
}

}  // namespace fpd
