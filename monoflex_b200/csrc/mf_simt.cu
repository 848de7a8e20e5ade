// CUDA-core cross-check of the tensor-core implicit GEMM (same operands, same epilogue, one thread per output).
// Debug/diagnostic only: selected with mf_set_conv_impl(1); never used by bench.py or the default path.
#include "mf_common.cuh"
#include "mf_kernels.h"

namespace mf {

__global__ void simt_gemm_kernel(const IgemmParams p, const __half* __restrict__ wp, int k_pad, int mode) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int ncols = (p.out_mode == OUT_F32_NHWC) ? p.y_ld : p.Cout;
  if (idx >= static_cast<long long>(p.M) * ncols) return;
  const int n = static_cast<int>(idx % ncols);
  const int m = static_cast<int>(idx / ncols);
  const int HoWo = p.Ho * p.Wo;
  const int b = m / HoWo, rem = m - b * HoWo;
  const int oy = rem / p.Wo, ox = rem - oy * p.Wo;
  float acc = 0.f;
  if (n < p.Cout) {
    const __half* w = wp + static_cast<long long>(n) * k_pad;
    const int ntap = p.kh * p.kw;
    for (int tap = 0; tap < ntap; ++tap) {
      const int ky = tap / p.kw, kx = tap - ky * p.kw;
      if (mode == MODE_CONV) {
        const int iy = oy * p.stride - p.pad + ky, ix = ox * p.stride - p.pad + kx;
        if (iy < 0 || iy >= p.H || ix < 0 || ix >= p.W) continue;
        const __half* xp = p.x + (static_cast<long long>(b * p.H + iy) * p.W + ix) * p.x_ld;
        for (int c = 0; c < p.Cin; ++c) acc += __half2float(xp[c]) * __half2float(w[tap * p.Cin + c]);
      } else {
        const float* om = p.offmask + static_cast<long long>(m) * p.om_ld;
        const float h_im = static_cast<float>(oy - 1 + ky) + om[2 * tap];
        const float w_im = static_cast<float>(ox - 1 + kx) + om[2 * tap + 1];
        const float mk = om[18 + tap];
        if (!(h_im > -1.f && w_im > -1.f && h_im < static_cast<float>(p.H) && w_im < static_cast<float>(p.W))) continue;
        const float hlf = floorf(h_im), wlf = floorf(w_im);
        const float lh = h_im - hlf, lw = w_im - wlf, hh = 1.f - lh, hw = 1.f - lw;
        const int hl = static_cast<int>(hlf), wl = static_cast<int>(wlf), hi = hl + 1, wi = wl + 1;
        const __half* xb = p.x + static_cast<long long>(b) * p.H * p.W * p.x_ld;
        for (int c = 0; c < p.Cin; ++c) {
          const float v1 = (hl >= 0 && wl >= 0) ? __half2float(xb[static_cast<long long>(hl * p.W + wl) * p.x_ld + c]) : 0.f;
          const float v2 = (hl >= 0 && wi <= p.W - 1) ? __half2float(xb[static_cast<long long>(hl * p.W + wi) * p.x_ld + c]) : 0.f;
          const float v3 = (hi <= p.H - 1 && wl >= 0) ? __half2float(xb[static_cast<long long>(hi * p.W + wl) * p.x_ld + c]) : 0.f;
          const float v4 = (hi <= p.H - 1 && wi <= p.W - 1) ? __half2float(xb[static_cast<long long>(hi * p.W + wi) * p.x_ld + c]) : 0.f;
          const float val = (hh * hw * v1 + hh * lw * v2 + lh * hw * v3 + lh * lw * v4) * mk;
          acc += __half2float(__float2half_rn(val)) * __half2float(w[tap * p.Cin + c]);
        }
      }
    }
  }
  float v = acc * p.scale[n] + p.shift[n];
  if (p.res != nullptr && n < p.Cout) v += __half2float(p.res[static_cast<long long>(m) * p.res_ld + n]);
  if (p.act == ACT_RELU) v = fmaxf(v, 0.f);
  else if (p.act == ACT_LEAKY) v = v > 0.f ? v : 0.01f * v;
  else if (p.act == ACT_OFFMASK && n >= 18) v = 1.f / (1.f + __expf(-v));
  if (p.out_mode == OUT_F16_NHWC) reinterpret_cast<__half*>(p.y)[static_cast<long long>(m) * p.y_ld + n] = __float2half_rn(v);
  else if (p.out_mode == OUT_F32_NHWC) reinterpret_cast<float*>(p.y)[static_cast<long long>(m) * p.y_ld + n] = v;
  else reinterpret_cast<float*>(p.y)[(static_cast<long long>(b) * p.y_ld + n) * HoWo + rem] = v;
}

int launch_simt_gemm(const IgemmParams& p, const __half* wp, int n_pad, int k_pad, int mode, cudaStream_t st) {
  (void)n_pad;
  const int ncols = (p.out_mode == OUT_F32_NHWC) ? p.y_ld : p.Cout;
  const long long total = static_cast<long long>(p.M) * ncols;
  const int threads = 256;
  const long long blocks = (total + threads - 1) / threads;
  simt_gemm_kernel<<<static_cast<unsigned>(blocks), threads, 0, st>>>(p, wp, k_pad, mode);
  return check_cuda(cudaGetLastError(), "simt gemm launch");
}

}  // namespace mf
