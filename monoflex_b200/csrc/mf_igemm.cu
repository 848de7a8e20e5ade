// Implicit-GEMM convolution and fused DCNv2 (bilinear-gather-then-contract) on the sm_100a tensor cores.
//
//   D[m, n] = sum_k A[m, k] * Wp[n, k]        m = output pixel (b, oy, ox), n = output channel,
//                                             k = (tap, cin) with cin fastest (NHWC activations)
//
// One CTA computes a 128 x BLOCK_N tile. Warp roles (warp-uniform dispatch, Blackwell anatomy):
//   warps [0, NPW)  A-tile producers, then epilogue. A rows are *gathered*:
//                     MODE_CONV  zero-padded conv taps via 16-byte cp.async (LDGSTS) into the 128B-swizzled
//                                K-major tile (the layout a SWIZZLE_128B TMA box would produce);
//                     MODE_DCN   per (pixel, tap): 4 neighbour NHWC vectors, fp32 bilinear blend * mask, fp16,
//                                st.shared (replaces the reference's im2col columns buffer,
//                                src/cuda/dcn_v2_im2col_cuda.cu:125-195 + SgemmBatched dcn_v2_cuda.cu:152-163).
//   warp NPW        weight (B operand) tiles by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) on the same mbarrier
//   warp NPW+1      TMEM allocation + single-thread tcgen05.mma issue (kind::f16, fp32 accumulators in TMEM)
// Epilogue: tcgen05.ld -> y = acc*scale[n] + shift[n] (+ residual) -> activation -> fp16 NHWC / fp32 NHWC / fp32 NCHW.
#include "mf_common.cuh"
#include "mf_kernels.h"

namespace mf {

static constexpr int BLOCK_M = 128;
static constexpr int BLOCK_K = 64;                       // fp16 elements = one 128-byte swizzle row
static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;    // 16 KB
static constexpr int LAG = 2;                            // cp.async groups kept in flight per producer thread

template <int BLOCK_N>
struct TileCfg {
  static constexpr int STAGES = BLOCK_N >= 128 ? 3 : 4;
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int SMEM = STAGES * (A_BYTES + B_BYTES) + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;
};

MF_DEVINL float act_apply(float v, int act, int n) {
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_LEAKY) return v > 0.f ? v : 0.01f * v;
  if (act == ACT_OFFMASK) return n >= 18 ? 1.f / (1.f + __expf(-v)) : v;
  return v;
}

template <int BLOCK_N, int MODE, int NPW>
__global__ void __launch_bounds__((NPW + 2) * 32)
igemm_kernel(const __grid_constant__ CUtensorMap tmap_w, const IgemmParams p) {
  using Cfg = TileCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int B_BYTES = Cfg::B_BYTES;
  constexpr int NPT = NPW * 32;                 // producer threads
  constexpr int RPP = NPT / 8;                  // rows per pass (8 lanes = 8 x 16B chunks per row)
  constexpr int PASSES = BLOCK_M / RPP;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_smem = smem;
  uint8_t* b_smem = smem + STAGES * A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * (A_BYTES + B_BYTES));
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int ntn = (p.Cout + BLOCK_N - 1) / BLOCK_N;
  const int n_tile = blockIdx.x % ntn;
  const int m_tile = blockIdx.x / ntn;
  const int m0 = m_tile * BLOCK_M;
  const int n0 = n_tile * BLOCK_N;
  const int nkb = p.nkb;
  // MODE_DCN: the 128 rows of a tile are an 8 x 16 pixel block (2-D locality for the bilinear gather: the block's
  // neighbourhood is ~30 KB of NHWC lines and stays L1-resident); MODE_CONV: 128 consecutive pixels.
  const int tiles_x = (p.W + 15) >> 4, tiles_y = (p.H + 7) >> 3;
  const int tile_b = m_tile / (tiles_x * tiles_y);
  const int tile_t = m_tile - tile_b * (tiles_x * tiles_y);
  const int tile_y0 = (tile_t / tiles_x) << 3, tile_x0 = (tile_t % tiles_x) << 4;

  if (warp == NPW && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], NPT + 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == NPW + 1) tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp < NPW) {
    // ============================================================ A producers
    const int tid = threadIdx.x;
    const int j = tid & 7;          // 16-byte chunk inside the 128-byte K row
    const int rsub = tid >> 3;
    const int HoWo = p.Ho * p.Wo;
    if (MODE == MODE_CONV) {
      int iy0[PASSES], ix0[PASSES];
      long long base[PASSES];
#pragma unroll
      for (int q = 0; q < PASSES; ++q) {
        const int m = m0 + q * RPP + rsub;
        if (m < p.M) {
          const int b = m / HoWo, rem = m - b * HoWo;
          const int oy = rem / p.Wo, ox = rem - oy * p.Wo;
          iy0[q] = oy * p.stride - p.pad;
          ix0[q] = ox * p.stride - p.pad;
          base[q] = (static_cast<long long>(b * p.H + iy0[q]) * p.W + ix0[q]) * p.x_ld;
        } else {
          iy0[q] = -100000; ix0[q] = -100000; base[q] = 0;   // never in bounds -> zero fill
        }
      }
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(&empty_bar[s], ((kb / STAGES) & 1) ^ 1);
        const int k = kb * BLOCK_K + j * 8;
        const int tap = k / p.Cin;
        const int c0 = k - tap * p.Cin;
        const int ky = tap / p.kw, kx = tap - ky * p.kw;
        const bool kvalid = k < p.K_real;
        const long long koff = static_cast<long long>(ky * p.W + kx) * p.x_ld + c0;
        const uint32_t a_stage = smem_u32(a_smem + s * A_BYTES);
#pragma unroll
        for (int q = 0; q < PASSES; ++q) {
          const int r = q * RPP + rsub;
          const int iy = iy0[q] + ky, ix = ix0[q] + kx;
          const bool ok = kvalid && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
          const __half* src = ok ? p.x + base[q] + koff : p.x;
          cp_async16(a_stage + sw128_off(r, j), src, ok ? 16u : 0u);
        }
        cp_async_commit();
        if (kb >= LAG) {
          cp_async_wait<LAG>();
          fence_proxy_async();
          mbar_arrive(&full_bar[(kb - LAG) % STAGES]);
        }
      }
      cp_async_wait<0>();
      fence_proxy_async();
      for (int kb = (nkb > LAG ? nkb - LAG : 0); kb < nkb; ++kb) mbar_arrive(&full_bar[kb % STAGES]);
    } else {
      // ---------------------------------------------------------- DCNv2 gather (3x3, stride 1, pad 1, dil 1, dg 1)
      int py[PASSES], px[PASSES];
      long long ibase[PASSES];      // element offset of image b
      long long obase[PASSES];      // offset/mask row
#pragma unroll
      for (int q = 0; q < PASSES; ++q) {
        const int r = q * RPP + rsub;
        const int yy = tile_y0 + (r >> 4), xx = tile_x0 + (r & 15);
        if (yy < p.H && xx < p.W && tile_b < p.B) {
          py[q] = yy;
          px[q] = xx;
          ibase[q] = static_cast<long long>(tile_b) * p.H * p.W * p.x_ld;
          obase[q] = (static_cast<long long>(tile_b * p.H + yy) * p.W + xx) * p.om_ld;
        } else {
          py[q] = -1; px[q] = 0; ibase[q] = 0; obase[q] = 0;
        }
      }
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(&empty_bar[s], ((kb / STAGES) & 1) ^ 1);
        const int k = kb * BLOCK_K + j * 8;
        const int tap = k / p.Cin;            // Cin % 64 == 0: every chunk of this K block shares the tap
        const int c0 = k - tap * p.Cin;
        const int ky = tap / 3, kx = tap - ky * 3;
        uint8_t* a_stage = a_smem + s * A_BYTES;
#pragma unroll
        for (int q = 0; q < PASSES; ++q) {
          const int r = q * RPP + rsub;
          uint4 out = make_uint4(0u, 0u, 0u, 0u);
          if (py[q] >= 0 && tap < 9) {
            const float* om = p.offmask + obase[q];
            const float off_h = __ldg(om + 2 * tap), off_w = __ldg(om + 2 * tap + 1), mk = __ldg(om + 18 + tap);
            const float h_im = static_cast<float>(py[q] - 1 + ky) + off_h;
            const float w_im = static_cast<float>(px[q] - 1 + kx) + off_w;
            if (h_im > -1.f && w_im > -1.f && h_im < static_cast<float>(p.H) && w_im < static_cast<float>(p.W)) {
              const float hlf = floorf(h_im), wlf = floorf(w_im);
              const float lh = h_im - hlf, lw = w_im - wlf, hh = 1.f - lh, hw = 1.f - lw;
              const int hl = static_cast<int>(hlf), wl = static_cast<int>(wlf);
              const int hh_i = hl + 1, wh_i = wl + 1;
              const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
              const __half* xb = p.x + ibase[q] + c0;
              const uint4 z = make_uint4(0u, 0u, 0u, 0u);
              const bool t = hl >= 0, bt = hh_i <= p.H - 1, l = wl >= 0, rt = wh_i <= p.W - 1;
              const uint4 v1 = (t && l) ? __ldg(reinterpret_cast<const uint4*>(xb + static_cast<long long>(hl * p.W + wl) * p.x_ld)) : z;
              const uint4 v2 = (t && rt) ? __ldg(reinterpret_cast<const uint4*>(xb + static_cast<long long>(hl * p.W + wh_i) * p.x_ld)) : z;
              const uint4 v3 = (bt && l) ? __ldg(reinterpret_cast<const uint4*>(xb + static_cast<long long>(hh_i * p.W + wl) * p.x_ld)) : z;
              const uint4 v4 = (bt && rt) ? __ldg(reinterpret_cast<const uint4*>(xb + static_cast<long long>(hh_i * p.W + wh_i) * p.x_ld)) : z;
              const __half2* h1 = reinterpret_cast<const __half2*>(&v1);
              const __half2* h2 = reinterpret_cast<const __half2*>(&v2);
              const __half2* h3 = reinterpret_cast<const __half2*>(&v3);
              const __half2* h4 = reinterpret_cast<const __half2*>(&v4);
              __half2 o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f1 = __half22float2(h1[e]), f2 = __half22float2(h2[e]);
                const float2 f3 = __half22float2(h3[e]), f4 = __half22float2(h4[e]);
                const float vx = (w1 * f1.x + w2 * f2.x + w3 * f3.x + w4 * f4.x) * mk;
                const float vy = (w1 * f1.y + w2 * f2.y + w3 * f3.y + w4 * f4.y) * mk;
                o[e] = __floats2half2_rn(vx, vy);
              }
              out = *reinterpret_cast<uint4*>(o);
            }
          }
          *reinterpret_cast<uint4*>(a_stage + sw128_off(r, j)) = out;
        }
        fence_proxy_async();
        mbar_arrive(&full_bar[s]);
      }
    }

    // ============================================================ epilogue (same warps)
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    int m = m0 + row;
    bool mvalid = m < p.M;
    if (MODE == MODE_DCN) {
      const int yy = tile_y0 + (row >> 4), xx = tile_x0 + (row & 15);
      mvalid = yy < p.H && xx < p.W && tile_b < p.B;
      m = (tile_b * p.H + yy) * p.W + xx;
    }
    constexpr int CHUNK = BLOCK_N >= 32 ? 32 : 16;
    constexpr int NCHUNK = BLOCK_N / CHUNK;
    constexpr int CGROUPS = NPW / 4;            // warps with the same TMEM quadrant split the column chunks
    for (int ch = warp >> 2; ch < NCHUNK; ch += CGROUPS) {
      uint32_t r[CHUNK];
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + ch * CHUNK;
      if constexpr (CHUNK == 32) tmem_ld32(taddr, r); else tmem_ld16(taddr, r);
      tmem_ld_wait();
      const int nb = n0 + ch * CHUNK;
      float v[CHUNK];
#pragma unroll
      for (int i = 0; i < CHUNK; ++i) v[i] = __uint_as_float(r[i]) * __ldg(p.scale + nb + i) + __ldg(p.shift + nb + i);
      if (p.res != nullptr && mvalid) {
        const __half* rp = p.res + static_cast<long long>(m) * p.res_ld + nb;
#pragma unroll
        for (int i = 0; i < CHUNK; i += 8) {
          if (nb + i < p.Cout) {
            const uint4 rv = __ldg(reinterpret_cast<const uint4*>(rp + i));
            const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __half22float2(rh[e]);
              v[i + 2 * e] += f.x;
              v[i + 2 * e + 1] += f.y;
            }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < CHUNK; ++i) v[i] = act_apply(v[i], p.act, nb + i);
      if (mvalid) {
        if (p.out_mode == OUT_F16_NHWC) {
          __half* yp = reinterpret_cast<__half*>(p.y) + static_cast<long long>(m) * p.y_ld + nb;
#pragma unroll
          for (int i = 0; i < CHUNK; i += 8) {
            if (nb + i < p.Cout) {
              __half2 o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) o[e] = __floats2half2_rn(v[i + 2 * e], v[i + 2 * e + 1]);
              *reinterpret_cast<uint4*>(yp + i) = *reinterpret_cast<uint4*>(o);
            }
          }
        } else if (p.out_mode == OUT_F32_NHWC) {
          float* yp = reinterpret_cast<float*>(p.y) + static_cast<long long>(m) * p.y_ld + nb;
#pragma unroll
          for (int i = 0; i < CHUNK; i += 4) {
            if (nb + i < p.y_ld) *reinterpret_cast<float4*>(yp + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          }
        } else {  // OUT_F32_NCHW: y[(b*y_ld + n)*HoWo + pix]
          const int b = m / HoWo, pix = m - b * HoWo;
          float* yp = reinterpret_cast<float*>(p.y) + (static_cast<long long>(b) * p.y_ld + nb) * HoWo + pix;
#pragma unroll
          for (int i = 0; i < CHUNK; ++i) {
            if (nb + i < p.Cout) yp[static_cast<long long>(i) * HoWo] = v[i];
          }
        }
      }
    }
  } else if (warp == NPW) {
    // ============================================================ weight tiles by TMA
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(&empty_bar[s], ((kb / STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], B_BYTES);
        tma_load_2d(smem_u32(b_smem + s * B_BYTES), &tmap_w, &full_bar[s], kb * BLOCK_K, n0);
      }
    }
  } else {
    // ============================================================ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(BLOCK_M, BLOCK_N);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(&full_bar[s], (kb / STAGES) & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(a_smem + s * A_BYTES);
        const uint32_t b_addr = smem_u32(b_smem + s * B_BYTES);
#pragma unroll
        for (int k4 = 0; k4 < BLOCK_K / 16; ++k4) {
          umma_f16(tmem_base, umma_desc_sw128(a_addr + k4 * 32), umma_desc_sw128(b_addr + k4 * 32), idesc,
                   (kb | k4) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);      // frees the smem stage once these MMAs have read it
      }
      umma_commit(tmem_full_bar);        // accumulator complete -> epilogue
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == NPW + 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || ptr == nullptr) {
      set_error("cuTensorMapEncodeTiled entry point unavailable");
      return nullptr;
    }
    fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

int g_tunable[16] = {0};   // [0] extra dynamic smem (bytes) for DCN CTAs, [1] same for conv CTAs

template <int BLOCK_N, int MODE, int NPW>
static int launch_cfg(const CUtensorMap& tm, const IgemmParams& p, cudaStream_t st) {
  using Cfg = TileCfg<BLOCK_N>;
  auto kern = igemm_kernel<BLOCK_N, MODE, NPW>;
  static int attr_smem = 0;
  int smem = Cfg::SMEM + g_tunable[MODE == MODE_DCN ? 0 : 1];
  if (smem > 227 * 1024) smem = 227 * 1024;
  if (smem > attr_smem) {
    if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "smem attr"))
      return -1;
    attr_smem = smem;
  }
  const int ntn = (p.Cout + BLOCK_N - 1) / BLOCK_N;
  const int ntm = MODE == MODE_DCN ? p.B * ((p.H + 7) / 8) * ((p.W + 15) / 16) : (p.M + BLOCK_M - 1) / BLOCK_M;
  kern<<<ntn * ntm, (NPW + 2) * 32, smem, st>>>(tm, p);
  return check_cuda(cudaGetLastError(), "igemm launch");
}

int igemm_block_n(int cout) {
  if (cout <= 16) return 16;
  if (cout <= 32) return 32;
  if (cout <= 64) return 64;
  return 128;
}

// wp: packed weights [n_pad, k_pad] fp16 (k_pad % 64 == 0, n_pad % block_n == 0)
int launch_igemm(const IgemmParams& p, const __half* wp, int n_pad, int k_pad, int mode, cudaStream_t st) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return -1;
  const int bn = igemm_block_n(p.Cout);
  if (n_pad % bn != 0 || k_pad % BLOCK_K != 0 || k_pad < p.nkb * BLOCK_K) {
    set_error("igemm: packed weight shape [%d,%d] incompatible with block_n=%d nkb=%d", n_pad, k_pad, bn, p.nkb);
    return -1;
  }
  if (mode == MODE_DCN && (p.Cin % 64 != 0 || p.kh != 3 || p.kw != 3 || p.stride != 1 || p.pad != 1)) {
    set_error("dcn igemm: only 3x3 s1 p1 with Cin %% 64 == 0 is built (got Cin=%d)", p.Cin);
    return -1;
  }
  if (mode == MODE_CONV && (p.Cin % 8 != 0 || p.x_ld % 8 != 0)) {
    set_error("conv igemm: Cin and pixel stride must be multiples of 8 (got %d, %d)", p.Cin, p.x_ld);
    return -1;
  }
  CUtensorMap tm;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(k_pad), static_cast<cuuint64_t>(n_pad)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(k_pad) * 2};
  cuuint32_t box[2] = {BLOCK_K, static_cast<cuuint32_t>(bn)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(wp), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
    return -1;
  }
#define MF_DISPATCH(BN)                                                                   \
  if (bn == BN) {                                                                         \
    if (mode == MODE_DCN) return launch_cfg<BN, MODE_DCN, 8>(tm, p, st);                  \
    return launch_cfg<BN, MODE_CONV, 4>(tm, p, st);                                       \
  }
  MF_DISPATCH(16)
  MF_DISPATCH(32)
  MF_DISPATCH(64)
  MF_DISPATCH(128)
#undef MF_DISPATCH
  set_error("igemm: unsupported block_n %d", bn);
  return -1;
}

}  // namespace mf
