// loss.cu -- fused Fast-Pose-Distillation loss + gradient in ONE pass over the heat-maps.
//
// Reference semantics (lib/core/loss.py:21-39 JointsMSELoss, lib/core/function.py:127-134 fpd_train):
//   L(out, ref) = 1/J * sum_j 0.5 * mean_{b,hw} ( w[b,j]*out - w[b,j]*ref )^2
//               = 0.5/(J*B*hw) * sum (w*(out-ref))^2                      (weights enter squared)
//   pose = sum_s L(out_s, target);  kd = sum_s L(out_s, teacher_last);  loss = (1-a)*pose + a*kd
//   dloss/dout_s = w^2 * ((1-a)*(out_s-target) + a*(out_s-teacher)) / (J*B*hw)
// The reference launches ~6 tiny kernels per joint per stack per term (~770 launches for the 4-stack
// student); here every stack's loss terms and gradient come out of one HBM pass, with a deterministic
// two-stage fp64 reduction (no atomics).
//
// Layouts: student / teacher heat-maps are the network-internal NHWC [B,h,w,J]; `target` is the caller's
// NCHW [B,J,h,w] tensor (lib/dataset/JointsDataset.py target layout), staged through shared memory.
#include "common.cuh"
#include "kernels.h"

namespace fpd {
namespace {

constexpr int kPix = 64;       // pixels per block
constexpr int kThreads = 256;
constexpr int kMaxStacks = 8;

struct StackPtrs {
  const float* out[kMaxStacks];
  float* grad[kMaxStacks];
};

__global__ void __launch_bounds__(kThreads)
fpd_loss_kernel(StackPtrs sp, int S, const float* __restrict__ target, const float* __restrict__ teacher,
                const float* __restrict__ tw, float alpha, float inv_norm, float grad_scale, int B, int J, int hw,
                double* __restrict__ part /*[nblocks][2]*/) {
  extern __shared__ float tgt[];  // [J][kPix+1]
  const int blocks_per_img = hw / kPix;
  const int b = blockIdx.x / blocks_per_img;
  const int p0 = (blockIdx.x % blocks_per_img) * kPix;
  for (int i = threadIdx.x; i < J * kPix; i += kThreads) {
    const int j = i / kPix, px = i % kPix;
    tgt[j * (kPix + 1) + px] = __ldg(target + ((int64_t)b * J + j) * hw + p0 + px);
  }
  __syncthreads();
  float pose = 0.f, kd = 0.f;
  const int64_t base = ((int64_t)b * hw + p0) * J;
  for (int i = threadIdx.x; i < J * kPix; i += kThreads) {
    const int px = i / J, j = i % J;
    const float w = __ldg(tw + b * J + j);
    const float w2 = w * w;
    const float g = tgt[j * (kPix + 1) + px];
    const float t = teacher ? __ldg(teacher + base + i) : 0.f;
    for (int s = 0; s < S; ++s) {
      const float o = __ldg(sp.out[s] + base + i);
      const float dp = o - g;
      const float dk = o - t;
      pose = fmaf(w2 * dp, dp, pose);
      if (teacher) kd = fmaf(w2 * dk, dk, kd);
      if (sp.grad[s]) {
        const float gr = teacher ? ((1.f - alpha) * dp + alpha * dk) : dp;
        sp.grad[s][base + i] = grad_scale * w2 * gr * inv_norm;
      }
    }
  }
  __shared__ double red[2][kThreads / 32];
  double dpose = warp_sum((double)pose), dkd = warp_sum((double)kd);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { red[0][wid] = dpose; red[1][wid] = dkd; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, c = 0.0;
    for (int i = 0; i < kThreads / 32; ++i) { a += red[0][i]; c += red[1][i]; }
    part[(size_t)blockIdx.x * 2 + 0] = a;
    part[(size_t)blockIdx.x * 2 + 1] = c;
  }
}

__global__ void fpd_loss_final_kernel(const double* __restrict__ part, int nblocks, double half_inv_norm,
                                      float alpha, int has_teacher, float* __restrict__ losses) {
  __shared__ double red[2][8];
  double a = 0.0, c = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x) { a += part[(size_t)i * 2]; c += part[(size_t)i * 2 + 1]; }
  a = warp_sum(a); c = warp_sum(c);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { red[0][wid] = a; red[1][wid] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double sa = 0.0, sc = 0.0;
    for (int i = 0; i < (int)(blockDim.x / 32); ++i) { sa += red[0][i]; sc += red[1][i]; }
    const double pose = sa * half_inv_norm, kd = sc * half_inv_norm;
    losses[0] = (float)pose;
    losses[1] = (float)kd;
    losses[2] = has_teacher ? (float)((1.0 - (double)alpha) * pose + (double)alpha * kd) : (float)pose;
  }
}

// single-term JointsMSELoss on NCHW tensors (the reference module's own signature)
__global__ void __launch_bounds__(kThreads)
joints_mse_kernel(const float* __restrict__ out, const float* __restrict__ target, const float* __restrict__ tw,
                  float inv_norm, float* __restrict__ grad, int64_t n, int hw, double* __restrict__ part) {
  float acc = 0.f;
  double dacc = 0.0;
  int inner = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float w = tw ? __ldg(tw + i / hw) : 1.f;
    const float d = __ldg(out + i) - __ldg(target + i);
    acc = fmaf(w * w * d, d, acc);
    if (grad) grad[i] = w * w * d * inv_norm;
    if (++inner == 32) { dacc += (double)acc; acc = 0.f; inner = 0; }
  }
  dacc += (double)acc;
  __shared__ double red[kThreads / 32];
  dacc = warp_sum(dacc);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) red[wid] = dacc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < kThreads / 32; ++i) s += red[i];
    part[(size_t)blockIdx.x * 2] = s;
    part[(size_t)blockIdx.x * 2 + 1] = 0.0;
  }
}

}  // namespace

size_t fpd_loss_workspace_bytes(int B, int J, int h, int w) {
  const int64_t hw = (int64_t)h * w;
  int64_t nblocks = (int64_t)B * ((hw + kPix - 1) / kPix);
  if (nblocks < 2048) nblocks = 2048;
  return (size_t)nblocks * 2 * sizeof(double);
}

int fpd_loss(const float* const* outs, int S, const float* target_nchw, const float* teacher_nhwc, const float* tw,
             float alpha, float* const* grads, float grad_scale, float* losses, int B, int J, int h, int w,
             void* workspace, size_t ws_bytes, cudaStream_t stream) {
  const int hw = h * w;
  FPD_REQUIRE(S >= 1 && S <= kMaxStacks, "fpd_loss: S=%d out of range [1,%d]", S, kMaxStacks);
  FPD_REQUIRE(hw % kPix == 0, "fpd_loss: h*w=%d must be a multiple of %d", hw, kPix);
  FPD_REQUIRE(J >= 1 && J <= 64, "fpd_loss: J=%d out of range", J);
  FPD_REQUIRE(tw != nullptr, "fpd_loss: target_weight is required (pass ones for use_target_weight=False)");
  const int nblocks = B * (hw / kPix);
  FPD_REQUIRE(ws_bytes >= (size_t)nblocks * 2 * sizeof(double), "fpd_loss: workspace too small");
  StackPtrs sp{};
  for (int s = 0; s < S; ++s) {
    sp.out[s] = outs[s];
    sp.grad[s] = grads ? grads[s] : nullptr;
  }
  const double norm = (double)J * (double)B * (double)hw;
  fpd_loss_kernel<<<nblocks, kThreads, (size_t)J * (kPix + 1) * sizeof(float), stream>>>(
      sp, S, target_nchw, teacher_nhwc, tw, alpha, (float)(1.0 / norm), grad_scale, B, J, hw, (double*)workspace);
  FPD_LAUNCH_CHECK();
  fpd_loss_final_kernel<<<1, 256, 0, stream>>>((const double*)workspace, nblocks, 0.5 / norm, alpha,
                                               teacher_nhwc != nullptr, losses);
  FPD_LAUNCH_CHECK();
  return FPD_OK;
}

int joints_mse(const float* out_nchw, const float* target_nchw, const float* tw, float* loss, float* grad_out,
               float, int B, int J, int hw, void* workspace, size_t ws_bytes, cudaStream_t stream) {
  const int64_t n = (int64_t)B * J * hw;
  int nblocks = (int)((n + kThreads * 8 - 1) / (kThreads * 8));
  if (nblocks > 2048) nblocks = 2048;
  if (nblocks < 1) nblocks = 1;
  FPD_REQUIRE(ws_bytes >= (size_t)nblocks * 2 * sizeof(double), "joints_mse: workspace too small");
  const double norm = (double)J * (double)B * (double)hw;
  joints_mse_kernel<<<nblocks, kThreads, 0, stream>>>(out_nchw, target_nchw, tw, (float)(1.0 / norm), grad_out, n, hw,
                                                      (double*)workspace);
  FPD_LAUNCH_CHECK();
  fpd_loss_final_kernel<<<1, 256, 0, stream>>>((const double*)workspace, nblocks, 0.5 / norm, 0.f, 0, loss);
  FPD_LAUNCH_CHECK();
  return FPD_OK;
}

}  // namespace fpd
