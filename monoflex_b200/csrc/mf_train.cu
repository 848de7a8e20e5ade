// Training-side HBM kernels (SURVEY §8 rows R11 backward, R13).
//
//  * adamw_arena_kernel: torch.optim.AdamW over ONE flat parameter / gradient / moment arena. The reference builds one
//    param group per tensor (solver/__init__.py:10-24, lr = BASE_LR or BASE_LR * BIAS_LR_FACTOR when "bias" is in the
//    name) => ~280 groups and ~1000 tiny kernels per step in eager torch; here each tensor is padded to a multiple of
//    ADAMW_CHUNK elements inside the arena and a per-chunk lr table selects the group's lr, so the whole step is one
//    launch that streams 28 B/parameter (read p, g, m, v; write p, m, v).
//  * focal_loss_backward_kernel: d/dpred of FocalLoss.forward (model/layers/focal_loss.py:35-55) scaled by a device
//    scalar (loss weight / clamp(num_pos, 1), detector_loss.py:276).
#include "mf_common.cuh"
#include "mf_launch.h"

namespace mf {

// ---------------------------------------------------------------- AdamW (torch/optim/adamw.py single-tensor semantics)
//   p   <- p * (1 - lr * wd)
//   m   <- m + (g - m) * (1 - beta1)                (torch: exp_avg.lerp_(grad, 1 - beta1))
//   v   <- v * beta2 + (1 - beta2) * g * g          (torch: exp_avg_sq.mul_(beta2).addcmul_(g, g, value = 1 - beta2))
//   p   <- p - (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
// g is first multiplied by grad_scale (1 / world_size when the arena holds an NCCL SUM; 1 otherwise).
__global__ void __launch_bounds__(256) adamw_arena_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                          float* __restrict__ m, float* __restrict__ v,
                                                          const float* __restrict__ chunk_lr, long long n_chunks,
                                                          float beta1, float beta2, float eps, float wd, float bc1,
                                                          float rsqrt_bc2, float grad_scale, float lr_scale) {
  pdl_wait();
  constexpr int VEC_PER_CHUNK = MF_ADAMW_CHUNK / 4;
  const long long n_vec = n_chunks * VEC_PER_CHUNK;
  const float omb1 = 1.f - beta1, omb2 = 1.f - beta2;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_vec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float lr = __ldg(chunk_lr + i / VEC_PER_CHUNK) * lr_scale;
    if (lr == 0.f) continue;                                   // padding-only chunk / frozen tensor
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = __ldcs(reinterpret_cast<const float4*>(g) + i);
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    const float decay = 1.f - lr * wd, step = lr / bc1;
    float* pe = reinterpret_cast<float*>(&pp);
    float* me = reinterpret_cast<float*>(&mm);
    float* ve = reinterpret_cast<float*>(&vv);
    const float* ge = reinterpret_cast<const float*>(&gg);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = ge[k] * grad_scale;
      float pk = pe[k] * decay;
      const float mk = me[k] + (gk - me[k]) * omb1;
      const float vk = ve[k] * beta2 + omb2 * gk * gk;
      const float denom = sqrtf(vk) * rsqrt_bc2 + eps;
      pk -= step * (mk / denom);
      pe[k] = pk; me[k] = mk; ve[k] = vk;
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
}

int launch_adamw_arena(float* p, const float* g, float* m, float* v, const float* chunk_lr, long long n_chunks,
                       float beta1, float beta2, float eps, float wd, long long step, float grad_scale, float lr_scale,
                       cudaStream_t st) {
  if (n_chunks <= 0) return 0;
  if (step < 1) { set_error("adamw: step must be >= 1 (1-based count of the update being applied)"); return -1; }
  // bias corrections in double like torch's Python scalars
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), static_cast<double>(step));
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), static_cast<double>(step));
  const long long n_vec = n_chunks * (MF_ADAMW_CHUNK / 4);
  long long blocks = (n_vec + 255) / 256;
  const long long cap = 148LL * 8 * 4;                          // a few waves; grid-stride for the rest
  if (blocks > cap) blocks = cap;
  (void)launch_k(adamw_arena_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, st, p, g, m, v, chunk_lr,
                 n_chunks, beta1, beta2, eps, wd, static_cast<float>(bc1), static_cast<float>(1.0 / sqrt(bc2)),
                 grad_scale, lr_scale);
  return check_cuda(cudaGetLastError(), "adamw_arena");
}

// ---------------------------------------------------------------- focal loss backward (layers/focal_loss.py:35-55)
// L = -sum_{t==1} log(p)(1-p)^2 - sum_{0<=t<1} log(1-p) p^2 (1-t)^4
// dL/dp = -( (1-p)^2 / p - 2 (1-p) log p )                          at t == 1
//       = -(1-t)^4 ( 2 p log(1-p) - p^2 / (1-p) )                   at 0 <= t < 1,   0 elsewhere (ignored pixels, t = -1)
__global__ void focal_loss_backward_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, long long n,
                                           const float* __restrict__ scale, float* __restrict__ grad) {
  pdl_wait();
  const float s = __ldg(scale);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float p = __ldg(pred + i), t = __ldg(tgt + i);
    float d = 0.f;
    if (t == 1.f) {
      const float q = 1.f - p;
      d = -(q * q / p - 2.f * q * logf(p));
    } else if (t < 1.f && t >= 0.f) {
      const float omt = 1.f - t, omt2 = omt * omt;
      d = -omt2 * omt2 * (2.f * p * logf(1.f - p) - p * p / (1.f - p));
    }
    grad[i] = d * s;
  }
}

int launch_focal_loss_backward(const float* pred, const float* tgt, long long n, const float* scale, float* grad,
                               cudaStream_t st) {
  if (n <= 0) return 0;
  long long blocks = (n + 1023) / 1024;
  if (blocks > 148 * 8) blocks = 148 * 8;
  (void)launch_k(focal_loss_backward_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, st, pred, tgt, n, scale,
                 grad);
  return check_cuda(cudaGetLastError(), "focal_loss_backward");
}

}  // namespace mf
