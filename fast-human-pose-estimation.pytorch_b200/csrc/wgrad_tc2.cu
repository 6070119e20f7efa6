// wgrad_tc2.cu -- convolution weight gradient on the sm_100a tensor cores, operand preparation fused in.
//
// Both operands are read RAW (fp32 x and dY, one TMA box each) and eight transform warps produce, in shared memory,
//     a = relu?((x - mean) * scale + shift)  (zeroed at the 3x3 padding positions)  ->  (a_hi, a_lo)
//     dY                                                                            ->  (g_hi, g_lo)
// before the MMA warp consumes the stage -- the backward pass then needs no separate operand-split passes at all.
//
//   dW[co][ci][tap] = scale * sum_{pixels p} dY[p][co] * A[p (+) tap][ci]        (stride 1, "same" padding)
//
// i.e. the `convolution_backward` weight half that dominates the reference's step (40 % of the
// CPU step; reference call site: loss.backward() in lib/core/function.py:146 through nn.Conv2d of
// lib/models/hourglass.py:20-27).
//
// GEMM view: the contraction runs over PIXELS, so both operands are read "transposed": a TMA box of
// [32 pixels x 32 channels] (128-byte rows, SWIZZLE_128B) is exactly the canonical *MN-major* UMMA operand
// tile (channels contiguous = the M/N axis, pixels = K, 8-pixel swizzle atoms). No transposition pass and no
// (tf32 MN-major operands must use the "128B swizzle with 32-byte atoms" layout: TMA SWIZZLE_128B_ATOM_32B <->
// UMMA layout type SWIZZLE_128B_BASE32B, 4-row atoms, SBO = 512 B.)
// im2col: a 3x3 tap is a shifted TMA box with hardware zero fill, and taps are *stacked along M*:
//   mode A (activation side on M): M = 128 = G taps x Cin (Cin in {32,64,128}, G = 128/Cin), N = Cout
//   mode B (dY side on M, 1x1 only): M = 128 output channels, N = Cin
// Work = (tap-group | M-tile) x split-K over pixel tiles; each CTA accumulates its K range in TMEM
// (tcgen05.mma kind::tf32, both operands MN-major), stores an fp32 partial, and a second kernel reduces the
// partials in a fixed order (deterministic) straight into the OIHW gradient tensor.
// 3xTF32 (hi/lo operand pairs) keeps fp32-grade accuracy, as in conv_tc.cu.
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace fpd {
namespace {

constexpr int kKp = 32;                // pixels (K) per pipeline stage
constexpr int kBoxBytes = kKp * 128;   // one [32 px x 32 ch] tf32 box
constexpr int kThreads = 448;
constexpr int kTransformThreads = 256;
constexpr int kMaxCin = 512;

struct WgradParams {
  int B, H, W, Cin, Cout, taps, passes;
  int mode_b;        // 0: activation side on M (tap stacking), 1: dY side on M
  int G;             // taps per group (mode A)
  int groups;        // tap groups (mode A) or M tiles (mode B)
  int N;             // UMMA N
  int nblk_n;        // N / 32
  int bn, bh, bw, tiles_w, tiles_h, tiles_n, num_ktiles;
  int splits, kt_per_split;
  int stages, stage_bytes, tmem_cols;
  float* partial;    // [splits][groups][128][N]
  const float* pre_mean;   // activation pre-op: a = relu?((x - mean) * scale + shift); null = identity
  const float* pre_scale;
  const float* pre_shift;
  int pre_relu;
  int dbg;   // timing ablations (FPD_WGRAD_DBG bit mask): 1 no MMA, 8 no transform work, 2 no TMA. Garbage results when set.
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
wgrad_tc_fused_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_g_hi,
                      const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // experiment (FPD_WGRAD_DBG & 16/32/64): every operand tile starts 1/2/3 rows (128 B each) off the swizzle-atom
  // boundary -- do the TMA write pattern and the UMMA read pattern stay consistent (both keyed on absolute address bits)?
  smem += ((p.dbg >> 4) & 3) * 128;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * p.stage_bytes);
  uint64_t* ready_bar = full_bar + p.stages;
  uint64_t* empty_bar = ready_bar + p.stages;
  uint64_t* done_bar = empty_bar + p.stages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(done_bar + 1);
  float* s_mean = reinterpret_cast<float*>(smem + (size_t)p.stages * p.stage_bytes + 256);
  float* s_scale = s_mean + kMaxCin;
  float* s_shift = s_scale + kMaxCin;

  // warp index made provably warp-uniform so the elected-lane TMA / MMA issue code keeps its descriptors in uniform
  // registers (a threadIdx-derived `if (lane == 0)` region makes ptxas wrap every UTCHMMA / UTMALDG in an
  // ELECT + R2UR.BROADCAST waterfall loop, ~100 cycles per MMA -- see conv_tc5.cu)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int g = blockIdx.x % p.groups;
  const int sp = blockIdx.x / p.groups;
  const int kt0 = sp * p.kt_per_split;
  const int kt1 = min(kt0 + p.kt_per_split, p.num_ktiles);
  const int nkt = max(kt1 - kt0, 0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a_hi);
    tma_prefetch_desc(&tm_g_hi);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&ready_bar[s], kTransformThreads / 32);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(done_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc_dyn(tmem_ptr_smem, (uint32_t)p.tmem_cols);
  {
    const int cin_pad = (p.Cin + 31) / 32 * 32;
    for (int c = threadIdx.x; c < cin_pad; c += kThreads) {
      const bool in = c < p.Cin;
      s_mean[c] = (in && p.pre_mean) ? p.pre_mean[c] : 0.f;
      s_scale[c] = (in && p.pre_scale) ? p.pre_scale[c] : 1.f;
      s_shift[c] = (in && p.pre_scale) ? p.pre_shift[c] : 0.f;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int half_bytes = (4 + p.nblk_n) * kBoxBytes;  // hi (or lo) part of a stage: 4 M boxes + N boxes
  const int cblks = (p.Cin + 31) / 32;
  const bool split = p.passes == 3;

  if (warp == 0) {
    {
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < nkt; ++i, s = (s + 1 == p.stages ? 0 : s + 1), ph ^= (s == 0)) {
        const int kt = kt0 + i;
        mbar_wait(&empty_bar[s], ph ^ 1);
        if (!elect_one()) { __syncwarp(); continue; }
        const int tw = kt % p.tiles_w, th = (kt / p.tiles_w) % p.tiles_h, tn = kt / (p.tiles_w * p.tiles_h);
        const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
        uint8_t* st = smem + (size_t)s * p.stage_bytes;
        if (p.dbg & 2) { mbar_arrive(&full_bar[s]); __syncwarp(); continue; }
        mbar_expect_tx(&full_bar[s], (uint32_t)half_bytes);   // raw operands only; the lo halves are produced on chip
        for (int j = 0; j < 4; ++j) {
          uint8_t* dst = st + j * kBoxBytes;
          if (!p.mode_b) {
            int tap = g * p.G + j / cblks;
            if (tap >= p.taps) tap = p.taps - 1;  // padded slot of the last group: result discarded
            const int cb = j % cblks;
            const int dh = (p.taps == 9) ? tap / 3 - 1 : 0, dw = (p.taps == 9) ? tap % 3 - 1 : 0;
            tma_load_4d(dst, &tm_a_hi, &full_bar[s], cb * 32, w0 + dw, h0 + dh, n0);
          } else {
            tma_load_4d(dst, &tm_g_hi, &full_bar[s], g * 128 + j * 32, w0, h0, n0);
          }
        }
        for (int j = 0; j < p.nblk_n; ++j) {
          uint8_t* dst = st + (4 + j) * kBoxBytes;
          tma_load_4d(dst, p.mode_b ? &tm_a_hi : &tm_g_hi, &full_bar[s], j * 32, w0, h0, n0);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    {
      const uint32_t idesc = umma_idesc_tf32(128, (uint32_t)p.N, 1, 1);  // both operands MN-major
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < nkt; ++i, s = (s + 1 == p.stages ? 0 : s + 1), ph ^= (s == 0)) {
        mbar_wait(&ready_bar[s], ph);   // transform done (implies the TMA bytes have landed)
        tc_fence_after_sync();
        if (!elect_one()) { __syncwarp(); continue; }
        const uint32_t m_hi = smem_u32(smem + (size_t)s * p.stage_bytes);
        const uint32_t n_hi = m_hi + 4 * kBoxBytes;
        const uint32_t m_lo = m_hi + half_bytes;
        const uint32_t n_lo = m_lo + 4 * kBoxBytes;
#pragma unroll
        for (int ks = 0; ks < kKp / 8; ++ks) {
          if (p.dbg & 1) break;
          const uint32_t koff = ks * 1024;  // 8 pixels = two 4-row (512 B) swizzle atoms; SBO = 512 steps between them
          const uint64_t dm_hi = umma_desc_sw128_32b(m_hi + koff, kBoxBytes, 512);
          const uint64_t dn_hi = umma_desc_sw128_32b(n_hi + koff, kBoxBytes, 512);
          uint32_t acc = (i > 0 || ks > 0) ? 1u : 0u;
          if (p.passes == 3) {
            const uint64_t dm_lo = umma_desc_sw128_32b(m_lo + koff, kBoxBytes, 512);
            const uint64_t dn_lo = umma_desc_sw128_32b(n_lo + koff, kBoxBytes, 512);
            umma_tf32(tmem_base, dm_lo, dn_hi, idesc, acc);
            umma_tf32(tmem_base, dm_hi, dn_lo, idesc, 1u);
            acc = 1u;
          }
          umma_tf32(tmem_base, dm_hi, dn_hi, idesc, acc);
        }
        umma_commit(&empty_bar[s]);
        __syncwarp();
      }
      if (elect_one()) umma_commit(done_bar);
      __syncwarp();
    }
  } else if (warp < 6) {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    float* prow = p.partial + (((size_t)sp * p.groups + g) * 128 + m) * p.N;
    if (nkt > 0) {
      mbar_wait(done_bar, 0);
      tc_fence_after_sync();
    }
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int c0 = 0; c0 < p.N; c0 += 16) {
      uint32_t v[16];
      if (nkt > 0) {
        tmem_ld16(taddr + c0, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0u;
      }
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        *reinterpret_cast<float4*>(prow + c0 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
    }
  }

  else {
    // ===================== operand transform =====================
    // thread t: pixel row r = t / 8 of every 32-pixel box, physical 16-byte chunk pc = t % 8 of that row. Under
    // SWIZZLE_128B_ATOM_32B the 32-byte chunk index is XORed with (row & 3): logical chunk = (((pc>>1)^(r&3))<<1)|(pc&1).
    const int t = threadIdx.x - 6 * 32;
    const int r = t >> 3, pc = t & 7;
    const int lc = ((((pc >> 1) ^ (r & 3)) << 1) | (pc & 1));   // logical 16-byte chunk -> channels 4*lc .. 4*lc+3
    const int dn = r / (p.bw * p.bh), dh_ = (r / p.bw) % p.bh, dw_ = r % p.bw;
    const bool has_affine = p.pre_scale != nullptr;
    const uint32_t row_off = (uint32_t)r * 128u + (uint32_t)pc * 16u;
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < nkt; ++i, s = (s + 1 == p.stages ? 0 : s + 1), ph ^= (s == 0)) {
      const int kt = kt0 + i;
      const int tw = kt % p.tiles_w, th = (kt / p.tiles_w) % p.tiles_h, tn = kt / (p.tiles_w * p.tiles_h);
      const int pn = tn * p.bn + dn, hh0 = th * p.bh + dh_, ww0 = tw * p.bw + dw_;
      mbar_wait(&full_bar[s], ph);
      const uint32_t base = smem_u32(smem + (size_t)s * p.stage_bytes);
      // three boxes per round: issue the raw (and parameter) loads of all three before the dependent arithmetic and
      // stores, so a stage costs ~2 load round trips instead of one per box (the serial form was the largest single
      // cost of this kernel: tools/diag_wgrad.py, 120 of 287 us on 3x3 64->64 @64x64)
      const int nbox = (p.dbg & 8) ? 0 : 4 + p.nblk_n;
      for (int j0 = 0; j0 < nbox; j0 += 3) {
        float4 v[3], mu[3], sc[3], sh[3];
        bool act[3], inb[3];
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const int j = j0 + u;
          act[u] = false; inb[u] = true;
          if (j >= nbox) continue;
          const uint32_t addr = base + (uint32_t)j * kBoxBytes + row_off;
          v[u] = lds128(addr);
          const bool is_act = (j < 4) ? !p.mode_b : (p.mode_b != 0);
          if (!is_act) continue;
          act[u] = true;
          int cb, hh = hh0, ww = ww0;
          if (j < 4) {   // mode A, M side: (tap, channel block)
            int tap = g * p.G + j / cblks;
            if (tap >= p.taps) tap = p.taps - 1;
            cb = j % cblks;
            if (p.taps == 9) { hh += tap / 3 - 1; ww += tap % 3 - 1; }
          } else {       // mode B, N side: channel block j-4, no tap shift (1x1)
            cb = j - 4;
          }
          const int c = cb * 32 + lc * 4;
          inb[u] = pn < p.B && hh >= 0 && hh < p.H && ww >= 0 && ww < p.W && c < p.Cin;
          if (inb[u] && has_affine) {
            mu[u] = lds128(smem_u32(s_mean) + (uint32_t)c * 4u);
            sc[u] = lds128(smem_u32(s_scale) + (uint32_t)c * 4u);
            sh[u] = lds128(smem_u32(s_shift) + (uint32_t)c * 4u);
          }
        }
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const int j = j0 + u;
          if (j >= nbox) continue;
          const uint32_t addr = base + (uint32_t)j * kBoxBytes + row_off;
          float4 x4 = v[u];
          if (act[u]) {
            if (inb[u]) {
              if (has_affine) {
                x4.x = fmaf(x4.x - mu[u].x, sc[u].x, sh[u].x); x4.y = fmaf(x4.y - mu[u].y, sc[u].y, sh[u].y);
                x4.z = fmaf(x4.z - mu[u].z, sc[u].z, sh[u].z); x4.w = fmaf(x4.w - mu[u].w, sc[u].w, sh[u].w);
              }
              if (p.pre_relu) {
                x4.x = fmaxf(x4.x, 0.f); x4.y = fmaxf(x4.y, 0.f); x4.z = fmaxf(x4.z, 0.f); x4.w = fmaxf(x4.w, 0.f);
              }
            } else {
              x4 = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
          float4 h, l;
          split_tf32_fast(x4.x, h.x, l.x); split_tf32_fast(x4.y, h.y, l.y);
          split_tf32_fast(x4.z, h.z, l.z); split_tf32_fast(x4.w, h.w, l.w);
          sts128(addr, h);
          if (split) sts128(addr + (uint32_t)half_bytes, l);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();                       // one arrival per warp: 256 per-thread arrivals on one barrier serialise
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

// dw[co][ci][tap] = scale * sum_splits partial[split][group][m][n]
// Block = 32 consecutive partial elements (coalesced along n) x 8 split groups: every thread sums splits sg, sg+8, ... and
// the eight group sums are added in a fixed order (deterministic). The one-thread-per-output form walked all <= 148
// splits serially (12 us per launch, latency-bound).
__global__ void __launch_bounds__(256)
wgrad_reduce2_kernel(const float* __restrict__ partial, float* __restrict__ dw, float scale, int Cin, int Cout,
                     int taps, int mode_b, int G, int groups, int N, int splits) {
  __shared__ float red[8][32];
  const int64_t per_split = (int64_t)groups * 128 * N;
  const int lane = threadIdx.x & 31, sg = threadIdx.x >> 5;
  const int64_t j = (int64_t)blockIdx.x * 32 + lane;   // index into [group][m][n]
  float acc = 0.f;
  if (j < per_split) {
#pragma unroll 4
    for (int s = sg; s < splits; s += 8) acc += partial[(size_t)s * per_split + j];
  }
  red[sg][lane] = acc;
  __syncthreads();
  if (sg != 0 || j >= per_split) return;
  float t = red[0][lane];
#pragma unroll
  for (int k = 1; k < 8; ++k) t += red[k][lane];
  const int n = (int)(j % N);
  const int m = (int)((j / N) % 128);
  const int g = (int)(j / ((int64_t)N * 128));
  int co, ci, tap;
  if (!mode_b) {
    tap = g * G + m / Cin;
    ci = m % Cin;
    co = n;
  } else {
    tap = 0;
    co = g * 128 + m;
    ci = n;
  }
  if (tap < taps && co < Cout && ci < Cin) dw[((size_t)co * Cin + ci) * taps + tap] = t * scale;
}

int pow2_div(int x, int cap) {
  int r = 1;
  while (r * 2 <= cap && x % (r * 2) == 0) r *= 2;
  return r;
}

struct Plan {
  bool ok;
  int mode_b, G, groups, N;
};

Plan make_plan(int Cin, int Cout, int ksize) {
  Plan pl{false, 0, 1, 1, 0};
  if (!(ksize == 1 || ksize == 3)) return pl;
  const int taps = ksize * ksize;
  if ((Cin == 32 || Cin == 64 || Cin == 128) && taps >= 128 / Cin && Cout % 32 == 0 && Cout >= 32 && Cout <= 256) {
    pl.ok = true; pl.mode_b = 0; pl.G = 128 / Cin; pl.groups = (taps + pl.G - 1) / pl.G; pl.N = Cout;
    return pl;
  }
  // mode B: dY channels on M in tiles of 128; channel counts that are not multiples of 128 / 32 are completed by the
  // TMA unit's out-of-bounds zero fill (costs no memory traffic), e.g. the 16-channel score convs or 32->64 1x1s.
  if (ksize == 1 && Cout % 4 == 0 && Cin % 4 == 0 && Cin >= 4 && Cin <= 256) {
    pl.ok = true; pl.mode_b = 1; pl.G = 1; pl.groups = (Cout + 127) / 128; pl.N = (Cin + 31) / 32 * 32;
    return pl;
  }
  return pl;
}

int choose_splits(int groups, int num_ktiles, int num_sms) {
  int s = num_sms / groups;
  if (s < 1) s = 1;
  if (s > num_ktiles) s = num_ktiles;
  return s;
}

}  // namespace

bool wgrad_tc_supported(int Cin, int Cout, int ksize) { return make_plan(Cin, Cout, ksize).ok; }

size_t wgrad_tc_workspace_bytes(int B, int H, int W, int Cin, int Cout, int ksize, int num_sms) {
  Plan pl = make_plan(Cin, Cout, ksize);
  if (!pl.ok) return 0;
  const int bw = pow2_div(W, kKp), bh = pow2_div(H, kKp / bw), bn = kKp / (bw * bh);
  const int num_ktiles = (W / bw) * (H / bh) * ((B + bn - 1) / bn);
  const int splits = choose_splits(pl.groups, num_ktiles, num_sms);
  return (size_t)splits * pl.groups * 128 * pl.N * sizeof(float);
}

int wgrad_tc_fused_launch(const float* x, const float* pre_mean, const float* pre_scale, const float* pre_shift,
                          int pre_relu, const float* dy, int passes, float* dw_oihw, float scale, int B, int H, int W,
                          int Cin, int Cout, int ksize, void* workspace, size_t workspace_bytes, int num_sms,
                          cudaStream_t stream) {
  if (wgrad_tc3_supported(H, W, Cin, Cout, ksize))   // 3x3: halo-tile kernel (csrc/wgrad_tc3.cu)
    return wgrad_tc3_launch(x, pre_mean, pre_scale, pre_shift, pre_relu, dy, passes, dw_oihw, scale, B, H, W, Cin, Cout,
                            workspace, workspace_bytes, num_sms, stream);
  const float* a_hi = x;
  const float* dy_hi = dy;
  Plan pl = make_plan(Cin, Cout, ksize);
  FPD_REQUIRE(pl.ok, "wgrad_tc: unsupported shape Cin=%d Cout=%d k=%d", Cin, Cout, ksize);
  FPD_REQUIRE(a_hi && dy_hi && dw_oihw, "wgrad_tc: null operand");
  FPD_REQUIRE(passes == 1 || passes == 3, "wgrad_tc_fused: passes must be 1 or 3");
  FPD_REQUIRE(Cin <= kMaxCin, "wgrad_tc_fused: Cin too large");
  FPD_REQUIRE((pre_scale == nullptr) == (pre_shift == nullptr), "wgrad_tc_fused: pre_scale/pre_shift come in pairs");
  WgradParams p{};
  p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.taps = ksize * ksize;
  p.passes = passes;
  p.pre_mean = pre_mean; p.pre_scale = pre_scale; p.pre_shift = pre_shift; p.pre_relu = pre_relu;
  {
    const char* e = getenv("FPD_WGRAD_DBG");
    p.dbg = e ? atoi(e) : 0;
  }
  p.mode_b = pl.mode_b; p.G = pl.G; p.groups = pl.groups; p.N = pl.N; p.nblk_n = pl.N / 32;
  p.bw = pow2_div(W, kKp);
  p.bh = pow2_div(H, kKp / p.bw);
  p.bn = kKp / (p.bw * p.bh);
  p.tiles_w = W / p.bw; p.tiles_h = H / p.bh; p.tiles_n = (B + p.bn - 1) / p.bn;
  p.num_ktiles = p.tiles_w * p.tiles_h * p.tiles_n;
  p.splits = choose_splits(p.groups, p.num_ktiles, num_sms);
  p.kt_per_split = (p.num_ktiles + p.splits - 1) / p.splits;
  p.stage_bytes = (p.passes == 3 ? 2 : 1) * (4 + p.nblk_n) * kBoxBytes;
  const int tail_bytes = 256 + 3 * kMaxCin * (int)sizeof(float);
  int stages = (208 * 1024 - tail_bytes) / p.stage_bytes;
  if (stages > 8) stages = 8;
  FPD_REQUIRE(stages >= 2, "wgrad_tc: stage does not fit (%d B)", p.stage_bytes);
  p.stages = stages;
  int tc = 32;
  while (tc < p.N) tc *= 2;
  p.tmem_cols = tc;
  const size_t need = (size_t)p.splits * p.groups * 128 * p.N * sizeof(float);
  FPD_REQUIRE(workspace && workspace_bytes >= need, "wgrad_tc: workspace too small (%zu < %zu)", workspace_bytes, need);
  p.partial = (float*)workspace;

  CUtensorMap tm_a_hi, tm_g_hi;
  uint32_t box[4] = {32u, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t strides[3] = {(uint64_t)Cin * 4, (uint64_t)W * Cin * 4, (uint64_t)H * W * Cin * 4};
    int rc = encode_tmap(&tm_a_hi, a_hi, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t strides[3] = {(uint64_t)Cout * 4, (uint64_t)W * Cout * 4, (uint64_t)H * W * Cout * 4};
    int rc = encode_tmap(&tm_g_hi, dy_hi, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    FPD_CUDA_CHECK(cudaFuncSetAttribute(wgrad_tc_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const size_t smem_bytes = (size_t)p.stages * p.stage_bytes + 1024 + tail_bytes + 512;
  wgrad_tc_fused_kernel<<<p.groups * p.splits, kThreads, smem_bytes, stream>>>(tm_a_hi, tm_g_hi, p);
  FPD_LAUNCH_CHECK();
  const int64_t per_split = (int64_t)p.groups * 128 * p.N;
  wgrad_reduce2_kernel<<<(int)((per_split + 31) / 32), 256, 0, stream>>>(p.partial, dw_oihw, scale, Cin, Cout, p.taps,
                                                                         p.mode_b, p.G, p.groups, p.N, p.splits);
  FPD_LAUNCH_CHECK();
  return FPD_OK;
}

}  // namespace fpd
