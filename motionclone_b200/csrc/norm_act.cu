// Memory-bound glue of the UNet3D forward on NHWC / token-major fp16 activations (sm_100a):
//   * GroupNorm(32) [+ SiLU] on channels_last `[(b f), h, w, C]`   (reference: models/resnet.py:21-29 InflatedGroupNorm,
//     :183-197; models/attention.py:61,105; models/motion_module.py:112,145 -- all via ATen GroupNorm on NCHW)
//   * LayerNorm over C                                             (models/attention.py:189-212, motion_module.py:204-215)
//   * GEGLU  h * gelu_erf(gate)                                    (diffusers-0.16 FeedForward used at attention.py:211,
//     motion_module.py:209)
// ATen's CUDA GroupNorm only takes NCHW: on a channels_last activation it costs a layout copy in, a layout copy before the
// next cuDNN conv, and two passes of its own. These kernels read NHWC directly with 128-bit accesses.
// Inference-only (no_grad) entry points; the autograd-carrying guided pass keeps ATen.
#include <math.h>

#include "mc_common.cuh"

namespace mc {

union Vec8 {
  uint4 u;
  __half2 h2[4];
  __half h[8];
};

// Chan et al. merge of (n, mean, M2) partials
__device__ __forceinline__ void welford_merge(float& n, float& mean, float& m2, float nb, float meanb, float m2b) {
  if (nb == 0.f) return;
  const float nn = n + nb;
  const float delta = meanb - mean;
  const float w = nb / nn;
  mean += delta * w;
  m2 += m2b + delta * delta * n * w;
  n = nn;
}

// ---------------------------------------------------------------------------------------------------------------
// GroupNorm, pass 1: per (frame, split) partial Welford statistics of every group.
// grid (N, S); CTA = 256 threads; thread -> (8-channel vector column v, pixel lane); partial[n][s][g] = {count, mean, M2}
// ---------------------------------------------------------------------------------------------------------------
constexpr int kGnMaxThreads = 512;  // 256 threads up to C = 2048, 512 up to C = 4096 (up-block concats reach 2560)

__global__ void __launch_bounds__(kGnMaxThreads) groupnorm_stats_kernel(const __half* __restrict__ x, float* __restrict__ partial,
                                                                    int HW, int C, int G, int S) {
  extern __shared__ float sm[];  // [C] sum, [C] sumsq-about-local-mean ... laid out as n, mean, m2 per channel
  const int n = blockIdx.x, s = blockIdx.y;
  const int V = C / 8;                       // vectors per pixel
  const int kGnThreads = blockDim.x;
  const int lanes = kGnThreads / V;          // pixel lanes handled concurrently (>= 1: V <= blockDim.x)
  const int v = threadIdx.x % V, pl = threadIdx.x / V;
  const int p_begin = (int)(((int64_t)HW * s) / S), p_end = (int)(((int64_t)HW * (s + 1)) / S);
  float sum[8], sq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sum[j] = sq[j] = 0.f;
  float cnt = 0.f;
  if (pl < lanes) {
    const __half* base = x + (int64_t)n * HW * C + v * 8;
    for (int p = p_begin + pl; p < p_end; p += lanes) {
      Vec8 a;
      a.u = *reinterpret_cast<const uint4*>(base + (int64_t)p * C);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float f = __half2float(a.h[j]);
        sum[j] += f;
        sq[j] += f * f;
      }
      cnt += 1.f;
    }
  }
  // per-thread (<= a few hundred samples per channel: fp32 sum / sum-of-squares is safe) -> (n, mean, M2) per channel,
  // then merged across pixel lanes and channels of a group with Chan's formula (robust to |mean| >> std)
  float* s_n = sm;
  float* s_mean = sm + kGnThreads * 8;
  float* s_m2 = sm + 2 * kGnThreads * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float mean = cnt > 0.f ? sum[j] / cnt : 0.f;
    const float m2 = cnt > 0.f ? fmaxf(sq[j] - sum[j] * mean, 0.f) : 0.f;
    const int slot = threadIdx.x * 8 + j;
    s_n[slot] = (pl < lanes) ? cnt : 0.f;
    s_mean[slot] = mean;
    s_m2[slot] = m2;
  }
  __syncthreads();
  const int cg = C / G;
  for (int g = threadIdx.x; g < G; g += kGnThreads) {
    float an = 0.f, amean = 0.f, am2 = 0.f;
    for (int c = g * cg; c < (g + 1) * cg; ++c) {
      const int vv = c / 8, jj = c % 8;
      for (int l = 0; l < lanes; ++l) {
        const int slot = (l * V + vv) * 8 + jj;
        welford_merge(an, amean, am2, s_n[slot], s_mean[slot], s_m2[slot]);
      }
    }
    float* out = partial + (((int64_t)n * S + s) * G + g) * 3;
    out[0] = an;
    out[1] = amean;
    out[2] = am2;
  }
}

// GroupNorm, pass 2: y = (x - mean) * rstd * gamma + beta [-> SiLU]; every CTA first folds the S partials of its frame.
template <bool SILU>
__global__ void __launch_bounds__(256) groupnorm_apply_kernel(const __half* __restrict__ x, __half* __restrict__ y,
                                                                    const float* __restrict__ partial,
                                                                    const __half* __restrict__ gamma,
                                                                    const __half* __restrict__ beta, int HW, int C, int G,
                                                                    int S, int chunks, float eps) {
  extern __shared__ float sm[];  // [G] mean, [G] rstd
  float* s_mean = sm;
  float* s_rstd = sm + G;
  constexpr int kGnThreads = 256;
  const int n = blockIdx.x, chunk = blockIdx.y;
  for (int g = threadIdx.x; g < G; g += kGnThreads) {
    float an = 0.f, amean = 0.f, am2 = 0.f;
    for (int s = 0; s < S; ++s) {
      const float* p = partial + (((int64_t)n * S + s) * G + g) * 3;
      welford_merge(an, amean, am2, p[0], p[1], p[2]);
    }
    s_mean[g] = amean;
    s_rstd[g] = rsqrtf(am2 / an + eps);
  }
  __syncthreads();
  const int V = C / 8, cg = C / G;
  const int64_t nvec = (int64_t)HW * V;
  const int64_t v_begin = nvec * chunk / chunks, v_end = nvec * (chunk + 1) / chunks;
  const __half* xb = x + (int64_t)n * HW * C;
  __half* yb = y + (int64_t)n * HW * C;
  for (int64_t i = v_begin + threadIdx.x; i < v_end; i += kGnThreads) {
    const int c0 = (int)(i % V) * 8;
    Vec8 a, w, b, o;
    a.u = *reinterpret_cast<const uint4*>(xb + i * 8);
    w.u = *reinterpret_cast<const uint4*>(gamma + c0);
    b.u = *reinterpret_cast<const uint4*>(beta + c0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (c0 + j) / cg;
      float f = (__half2float(a.h[j]) - s_mean[g]) * s_rstd[g] * __half2float(w.h[j]) + __half2float(b.h[j]);
      if (SILU) {
        f = round_half(f);             // ATen rounds the GroupNorm output to fp16 before the separate SiLU kernel
        f = f / (1.f + __expf(-f));
      }
      o.h[j] = __float2half_rn(f);
    }
    *reinterpret_cast<uint4*>(yb + i * 8) = o.u;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm over the last dim C (multiple of 8, <= 2048): one warp per row, two-pass statistics in registers
// ---------------------------------------------------------------------------------------------------------------
template <int VPL>  // vectors (of 8 halfs) per lane
__global__ void __launch_bounds__(256) layernorm_kernel(const __half* __restrict__ x, __half* __restrict__ y,
                                                        const __half* __restrict__ gamma, const __half* __restrict__ beta,
                                                        int64_t rows, int C, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int V = C / 8;
  for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < rows; row += (int64_t)gridDim.x * 8) {
    const __half* xr = x + row * C;
    Vec8 a[VPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v = lane + i * 32;
      if (v < V) {
        a[i].u = *reinterpret_cast<const uint4*>(xr + v * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += __half2float(a[i].h[j]);
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    const float mean = s / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v = lane + i * 32;
      if (v < V) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = __half2float(a[i].h[j]) - mean;
          q += d * d;
        }
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) q += __shfl_xor_sync(0xffffffffu, q, off);
    const float rstd = rsqrtf(q / (float)C + eps);
    __half* yr = y + row * C;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v = lane + i * 32;
      if (v < V) {
        Vec8 w, b, o;
        w.u = *reinterpret_cast<const uint4*>(gamma + v * 8);
        b.u = *reinterpret_cast<const uint4*>(beta + v * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          o.h[j] = __float2half_rn((__half2float(a[i].h[j]) - mean) * rstd * __half2float(w.h[j]) + __half2float(b.h[j]));
        *reinterpret_cast<uint4*>(yr + v * 8) = o.u;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// GEGLU: out[t, j] = fp16( h[t, j] * fp16(gelu_erf(gate[t, j])) ),  in = [T, 2I] = [h | gate]
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
}

__global__ void __launch_bounds__(256) geglu_kernel(const __half* __restrict__ in, __half* __restrict__ out, int64_t T,
                                                    int I) {
  const int VI = I / 8;
  const int64_t nvec = T * VI;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = i / VI;
    const int v = (int)(i % VI);
    const __half* row = in + t * 2 * I;
    Vec8 hh, gg, o;
    hh.u = *reinterpret_cast<const uint4*>(row + v * 8);
    gg.u = *reinterpret_cast<const uint4*>(row + I + v * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float ge = round_half(gelu_erf(__half2float(gg.h[j])));  // F.gelu output is fp16 in the eager graph
      o.h[j] = __float2half_rn(__half2float(hh.h[j]) * ge);
    }
    *reinterpret_cast<uint4*>(out + i * 8) = o.u;
  }
}

}  // namespace mc

extern "C" int mc_groupnorm_nhwc(const void* x, void* y, const void* gamma, const void* beta, void* workspace,
                                 int64_t workspace_bytes, int N, int HW, int C, int G, float eps, int fuse_silu,
                                 void* stream) {
  using namespace mc;
  if (!x || !y || !gamma || !beta || !workspace || N <= 0 || HW <= 0) {
    set_error("groupnorm_nhwc: null pointer or non-positive dims");
    return MC_E_INVALID;
  }
  if (C % 8 != 0 || C % G != 0 || C / 8 > kGnMaxThreads || G > 256) {
    set_error("groupnorm_nhwc: need C %% 8 == 0, C %% G == 0, C <= 4096, G <= 256 (got C=%d G=%d)", C, G);
    return MC_E_UNSUPPORTED;
  }
  // splits: enough CTAs to cover the machine (148 SMs x ~4) without shrinking a split below 64 pixels
  int S = (148 * 4 + N - 1) / N;
  if (S > HW / 64) S = HW / 64;
  if (S < 1) S = 1;
  if (S > 64) S = 64;
  const int64_t need = (int64_t)N * S * G * 3 * sizeof(float);
  if (workspace_bytes < need) {
    set_error("groupnorm_nhwc: workspace too small (%lld < %lld bytes)", (long long)workspace_bytes, (long long)need);
    return MC_E_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int threads1 = (C / 8 <= 256) ? 256 : 512;
  const int smem1 = 3 * threads1 * 8 * sizeof(float);  // 24 KB / 48 KB
  groupnorm_stats_kernel<<<dim3(N, S), threads1, smem1, st>>>((const __half*)x, (float*)workspace, HW, C, G, S);
  count_launch();
  int rc = check_launch("groupnorm_stats");
  if (rc != MC_OK) return rc;
  constexpr int kGnThreads = 256;
  int chunks = (148 * 8 + N - 1) / N;
  const int64_t nvec = (int64_t)HW * (C / 8);
  if (chunks > nvec / kGnThreads) chunks = (int)(nvec / kGnThreads);
  if (chunks < 1) chunks = 1;
  const int smem2 = 2 * G * sizeof(float);
  if (fuse_silu)
    groupnorm_apply_kernel<true><<<dim3(N, chunks), kGnThreads, smem2, st>>>(
        (const __half*)x, (__half*)y, (const float*)workspace, (const __half*)gamma, (const __half*)beta, HW, C, G, S,
        chunks, eps);
  else
    groupnorm_apply_kernel<false><<<dim3(N, chunks), kGnThreads, smem2, st>>>(
        (const __half*)x, (__half*)y, (const float*)workspace, (const __half*)gamma, (const __half*)beta, HW, C, G, S,
        chunks, eps);
  count_launch();
  return check_launch("groupnorm_apply");
}

extern "C" int64_t mc_groupnorm_workspace_bytes(int N, int G) { return (int64_t)N * 64 * G * 3 * sizeof(float); }

extern "C" int mc_layernorm(const void* x, void* y, const void* gamma, const void* beta, int64_t rows, int C, float eps,
                            void* stream) {
  using namespace mc;
  if (!x || !y || !gamma || !beta || rows <= 0) {
    set_error("layernorm: null pointer or rows <= 0");
    return MC_E_INVALID;
  }
  if (C % 8 != 0 || C > 2048) {
    set_error("layernorm: need C %% 8 == 0 and C <= 2048 (got %d)", C);
    return MC_E_UNSUPPORTED;
  }
  int64_t blocks = (rows + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cudaStream_t st = (cudaStream_t)stream;
  const int vpl = (C / 8 + 31) / 32;
#define MC_LN(V)                                                                                                \
  layernorm_kernel<V><<<(unsigned)blocks, 256, 0, st>>>((const __half*)x, (__half*)y, (const __half*)gamma,     \
                                                        (const __half*)beta, rows, C, eps)
  switch (vpl) {
    case 1: MC_LN(1); break;
    case 2: MC_LN(2); break;
    case 3: MC_LN(3); break;
    case 4: MC_LN(4); break;
    case 5: MC_LN(5); break;
    case 6: MC_LN(6); break;
    case 7: MC_LN(7); break;
    default: MC_LN(8); break;
  }
#undef MC_LN
  count_launch();
  return check_launch("layernorm");
}

extern "C" int mc_geglu(const void* in, void* out, int64_t T, int I, void* stream) {
  using namespace mc;
  if (!in || !out || T <= 0 || I <= 0) {
    set_error("geglu: null pointer or non-positive dims");
    return MC_E_INVALID;
  }
  if (I % 8 != 0) {
    set_error("geglu: inner dim must be a multiple of 8 (got %d)", I);
    return MC_E_UNSUPPORTED;
  }
  const int64_t nvec = T * (I / 8);
  int64_t blocks = (nvec + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  geglu_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const __half*)in, (__half*)out, T, I);
  count_launch();
  return check_launch("geglu");
}
