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

__global__ void __launch_bounds__(kGnMaxThreads) groupnorm_stats_kernel(const __half* __restrict__ x,
                                                                       const __half* __restrict__ chan_bias, int frames_per_row,
                                                                       float* __restrict__ partial, int HW, int C, int G,
                                                                       int S) {
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
    Vec8 cb;
    cb.u = make_uint4(0u, 0u, 0u, 0u);
    if (chan_bias != nullptr) cb.u = *reinterpret_cast<const uint4*>(chan_bias + (int64_t)(n / frames_per_row) * C + v * 8);
    for (int p = p_begin + pl; p < p_end; p += lanes) {
      Vec8 a;
      a.u = *reinterpret_cast<const uint4*>(base + (int64_t)p * C);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float f = __half2float(a.h[j]);
        if (chan_bias != nullptr) f = round_half(f + __half2float(cb.h[j]));  // the eager `h + temb` is an fp16 tensor
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
                                                                    const __half* __restrict__ chan_bias,
                                                                    int frames_per_row, const float* __restrict__ partial,
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
    Vec8 a, w, b, o, cb;
    a.u = *reinterpret_cast<const uint4*>(xb + i * 8);
    w.u = *reinterpret_cast<const uint4*>(gamma + c0);
    b.u = *reinterpret_cast<const uint4*>(beta + c0);
    if (chan_bias != nullptr) cb.u = *reinterpret_cast<const uint4*>(chan_bias + (int64_t)(n / frames_per_row) * C + c0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (c0 + j) / cg;
      float xv = __half2float(a.h[j]);
      if (chan_bias != nullptr) xv = round_half(xv + __half2float(cb.h[j]));
      float f = (xv - s_mean[g]) * s_rstd[g] * __half2float(w.h[j]) + __half2float(b.h[j]);
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
                                                        const __half* __restrict__ post_add, int rows_per_frame, int frames,
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
    const __half* pa = post_add ? post_add + (int64_t)((row / rows_per_frame) % frames) * C : nullptr;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v = lane + i * 32;
      if (v < V) {
        Vec8 w, b, o, pe;
        w.u = *reinterpret_cast<const uint4*>(gamma + v * 8);
        b.u = *reinterpret_cast<const uint4*>(beta + v * 8);
        if (pa) pe.u = *reinterpret_cast<const uint4*>(pa + v * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float f = (__half2float(a[i].h[j]) - mean) * rstd * __half2float(w.h[j]) + __half2float(b.h[j]);
          if (pa) f = round_half(f) + __half2float(pe.h[j]);  // eager: LayerNorm output (fp16) + pe (fp16)
          o.h[j] = __float2half_rn(f);
        }
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

extern "C" int mc_groupnorm_nhwc(const void* x, const void* chan_bias, int frames_per_bias_row, void* y,
                                 const void* gamma, const void* beta, void* workspace, int64_t workspace_bytes, int N,
                                 int HW, int C, int G, float eps, int fuse_silu, void* stream) {
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
  if (S > 64) S = 64;  // keep in sync with gn_splits() below
  const int64_t need = (int64_t)N * S * G * 3 * sizeof(float);
  if (workspace_bytes < need) {
    set_error("groupnorm_nhwc: workspace too small (%lld < %lld bytes)", (long long)workspace_bytes, (long long)need);
    return MC_E_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int threads1 = (C / 8 <= 256) ? 256 : 512;
  const int smem1 = 3 * threads1 * 8 * sizeof(float);  // 24 KB / 48 KB
  if (chan_bias != nullptr && frames_per_bias_row <= 0) {
    set_error("groupnorm_nhwc: frames_per_bias_row must be positive when chan_bias is given");
    return MC_E_INVALID;
  }
  groupnorm_stats_kernel<<<dim3(N, S), threads1, smem1, st>>>((const __half*)x, (const __half*)chan_bias,
                                                              frames_per_bias_row, (float*)workspace, HW, C, G, S);
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
        (const __half*)x, (__half*)y, (const __half*)chan_bias, frames_per_bias_row, (const float*)workspace,
        (const __half*)gamma, (const __half*)beta, HW, C, G, S, chunks, eps);
  else
    groupnorm_apply_kernel<false><<<dim3(N, chunks), kGnThreads, smem2, st>>>(
        (const __half*)x, (__half*)y, (const __half*)chan_bias, frames_per_bias_row, (const float*)workspace,
        (const __half*)gamma, (const __half*)beta, HW, C, G, S, chunks, eps);
  count_launch();
  return check_launch("groupnorm_apply");
}

extern "C" int64_t mc_groupnorm_workspace_bytes(int N, int G) { return (int64_t)N * 64 * G * 3 * sizeof(float); }

extern "C" int mc_layernorm(const void* x, void* y, const void* gamma, const void* beta, const void* post_add,
                            int rows_per_frame, int frames, int64_t rows, int C, float eps, void* stream) {
  using namespace mc;
  if (!x || !y || !gamma || !beta || rows <= 0) {
    set_error("layernorm: null pointer or rows <= 0");
    return MC_E_INVALID;
  }
  if (C % 8 != 0 || C > 2048) {
    set_error("layernorm: need C %% 8 == 0 and C <= 2048 (got %d)", C);
    return MC_E_UNSUPPORTED;
  }
  if (post_add != nullptr && (rows_per_frame <= 0 || frames <= 0)) {
    set_error("layernorm: post_add needs rows_per_frame > 0 and frames > 0");
    return MC_E_INVALID;
  }
  int64_t blocks = (rows + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cudaStream_t st = (cudaStream_t)stream;
  const int vpl = (C / 8 + 31) / 32;
#define MC_LN(V)                                                                                                \
  layernorm_kernel<V><<<(unsigned)blocks, 256, 0, st>>>((const __half*)x, (__half*)y, (const __half*)gamma,     \
                                                        (const __half*)beta, (const __half*)post_add, rows_per_frame, \
                                                        frames, rows, C, eps)
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

// ================================================================================================================
// Backward kernels for the autograd-carrying (guided, conditional) pass. Weights are frozen on this path
// (t2v_video_sample.py:67-68), so only input gradients are produced.
// ================================================================================================================
namespace mc {

__device__ __forceinline__ float silu_grad(float y) {
  const float s = 1.f / (1.f + __expf(-y));
  return s * (1.f + y * (1.f - s));
}

// GroupNorm(+SiLU) backward, pass 1: per (frame, split) partial sums over each group of
//   A = sum(dxhat), B = sum(dxhat * xhat), with dxhat = dy * gamma, dy = dz * silu'(y) when SiLU was fused.
template <bool SILU>
__global__ void __launch_bounds__(kGnMaxThreads) groupnorm_bwd_reduce_kernel(
    const __half* __restrict__ x, const __half* __restrict__ chan_bias, int frames_per_row, const __half* __restrict__ dz,
    const float* __restrict__ stats, const __half* __restrict__ gamma, const __half* __restrict__ beta,
    float* __restrict__ partial, int HW, int C, int G, int S) {
  extern __shared__ float sm[];
  const int n = blockIdx.x, s = blockIdx.y;
  const int V = C / 8, nthreads = blockDim.x, lanes = nthreads / V, cg = C / G;
  const int v = threadIdx.x % V, pl = threadIdx.x / V;
  const int p_begin = (int)(((int64_t)HW * s) / S), p_end = (int)(((int64_t)HW * (s + 1)) / S);
  float sa[8], sb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sa[j] = sb[j] = 0.f;
  if (pl < lanes) {
    const int c0 = v * 8;
    Vec8 w, b, cb;
    w.u = *reinterpret_cast<const uint4*>(gamma + c0);
    b.u = *reinterpret_cast<const uint4*>(beta + c0);
    cb.u = make_uint4(0u, 0u, 0u, 0u);
    if (chan_bias != nullptr) cb.u = *reinterpret_cast<const uint4*>(chan_bias + (int64_t)(n / frames_per_row) * C + c0);
    float mean[8], rstd[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (c0 + j) / cg;
      mean[j] = stats[((int64_t)n * G + g) * 2];
      rstd[j] = stats[((int64_t)n * G + g) * 2 + 1];
    }
    const __half* xb = x + (int64_t)n * HW * C + c0;
    const __half* db = dz + (int64_t)n * HW * C + c0;
    for (int p = p_begin + pl; p < p_end; p += lanes) {
      Vec8 a, d;
      a.u = *reinterpret_cast<const uint4*>(xb + (int64_t)p * C);
      d.u = *reinterpret_cast<const uint4*>(db + (int64_t)p * C);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float xv = __half2float(a.h[j]);
        if (chan_bias != nullptr) xv = round_half(xv + __half2float(cb.h[j]));
        const float xh = (xv - mean[j]) * rstd[j];
        float dy = __half2float(d.h[j]);
        if (SILU) dy *= silu_grad(round_half(xh * __half2float(w.h[j]) + __half2float(b.h[j])));
        const float dxh = dy * __half2float(w.h[j]);
        sa[j] += dxh;
        sb[j] += dxh * xh;
      }
    }
  }
  float* s_a = sm;
  float* s_b = sm + nthreads * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s_a[threadIdx.x * 8 + j] = (pl < lanes) ? sa[j] : 0.f;
    s_b[threadIdx.x * 8 + j] = (pl < lanes) ? sb[j] : 0.f;
  }
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += nthreads) {
    float ta = 0.f, tb = 0.f;
    for (int c = g * cg; c < (g + 1) * cg; ++c) {
      const int vv = c / 8, jj = c % 8;
      for (int l = 0; l < lanes; ++l) {
        ta += s_a[(l * V + vv) * 8 + jj];
        tb += s_b[(l * V + vv) * 8 + jj];
      }
    }
    float* out = partial + (((int64_t)n * S + s) * G + g) * 2;
    out[0] = ta;
    out[1] = tb;
  }
}

// pass 2: dx = rstd * (dxhat - A/m - xhat * B/m)
template <bool SILU>
__global__ void __launch_bounds__(256) groupnorm_bwd_apply_kernel(
    const __half* __restrict__ x, const __half* __restrict__ chan_bias, int frames_per_row, const __half* __restrict__ dz,
    __half* __restrict__ dx, const float* __restrict__ stats, const float* __restrict__ partial,
    const __half* __restrict__ gamma, const __half* __restrict__ beta, int HW, int C, int G, int S, int chunks) {
  extern __shared__ float sm[];  // [G] mean, rstd, A/m, B/m
  float* s_mean = sm;
  float* s_rstd = sm + G;
  float* s_a = sm + 2 * G;
  float* s_b = sm + 3 * G;
  const int n = blockIdx.x, chunk = blockIdx.y;
  const int V = C / 8, cg = C / G;
  const float inv_m = 1.f / ((float)HW * (float)cg);
  for (int g = threadIdx.x; g < G; g += 256) {
    float ta = 0.f, tb = 0.f;
    for (int s = 0; s < S; ++s) {
      const float* p = partial + (((int64_t)n * S + s) * G + g) * 2;
      ta += p[0];
      tb += p[1];
    }
    s_mean[g] = stats[((int64_t)n * G + g) * 2];
    s_rstd[g] = stats[((int64_t)n * G + g) * 2 + 1];
    s_a[g] = ta * inv_m;
    s_b[g] = tb * inv_m;
  }
  __syncthreads();
  const int64_t nvec = (int64_t)HW * V;
  const int64_t v_begin = nvec * chunk / chunks, v_end = nvec * (chunk + 1) / chunks;
  const __half* xb = x + (int64_t)n * HW * C;
  const __half* db = dz + (int64_t)n * HW * C;
  __half* ob = dx + (int64_t)n * HW * C;
  for (int64_t i = v_begin + threadIdx.x; i < v_end; i += 256) {
    const int c0 = (int)(i % V) * 8;
    Vec8 a, d, w, b, o, cb;
    a.u = *reinterpret_cast<const uint4*>(xb + i * 8);
    d.u = *reinterpret_cast<const uint4*>(db + i * 8);
    w.u = *reinterpret_cast<const uint4*>(gamma + c0);
    b.u = *reinterpret_cast<const uint4*>(beta + c0);
    if (chan_bias != nullptr) cb.u = *reinterpret_cast<const uint4*>(chan_bias + (int64_t)(n / frames_per_row) * C + c0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (c0 + j) / cg;
      float xv = __half2float(a.h[j]);
      if (chan_bias != nullptr) xv = round_half(xv + __half2float(cb.h[j]));
      const float xh = (xv - s_mean[g]) * s_rstd[g];
      float dy = __half2float(d.h[j]);
      if (SILU) dy *= silu_grad(round_half(xh * __half2float(w.h[j]) + __half2float(b.h[j])));
      const float dxh = dy * __half2float(w.h[j]);
      o.h[j] = __float2half_rn(s_rstd[g] * (dxh - s_a[g] - xh * s_b[g]));
    }
    *reinterpret_cast<uint4*>(ob + i * 8) = o.u;
  }
}

// finalise forward statistics into [N, G, 2] = (mean, rstd) for the backward
__global__ void groupnorm_finalize_stats_kernel(const float* __restrict__ partial, float* __restrict__ stats, int G, int S,
                                                float eps, int total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int n = i / G, g = i % G;
  float an = 0.f, amean = 0.f, am2 = 0.f;
  for (int s = 0; s < S; ++s) {
    const float* p = partial + (((int64_t)n * S + s) * G + g) * 3;
    welford_merge(an, amean, am2, p[0], p[1], p[2]);
  }
  stats[2 * i] = amean;
  stats[2 * i + 1] = rsqrtf(am2 / an + eps);
}

// LayerNorm backward (input gradient only): one warp per row, statistics recomputed from x
template <int VPL>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const __half* __restrict__ x, const __half* __restrict__ dy,
                                                            __half* __restrict__ dx, const __half* __restrict__ gamma,
                                                            int64_t rows, int C, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int V = C / 8;
  for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < rows; row += (int64_t)gridDim.x * 8) {
    const __half* xr = x + row * C;
    const __half* dr = dy + row * C;
    Vec8 a[VPL], d[VPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v = lane + i * 32;
      if (v < V) {
        a[i].u = *reinterpret_cast<const uint4*>(xr + v * 8);
        d[i].u = *reinterpret_cast<const uint4*>(dr + v * 8);
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
          const float t = __half2float(a[i].h[j]) - mean;
          q += t * t;
        }
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) q += __shfl_xor_sync(0xffffffffu, q, off);
    const float rstd = rsqrtf(q / (float)C + eps);
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v = lane + i * 32;
      if (v < V) {
        Vec8 w;
        w.u = *reinterpret_cast<const uint4*>(gamma + v * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float dxh = __half2float(d[i].h[j]) * __half2float(w.h[j]);
          sa += dxh;
          sb += dxh * (__half2float(a[i].h[j]) - mean) * rstd;
        }
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      sa += __shfl_xor_sync(0xffffffffu, sa, off);
      sb += __shfl_xor_sync(0xffffffffu, sb, off);
    }
    sa /= (float)C;
    sb /= (float)C;
    __half* outr = dx + row * C;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v = lane + i * 32;
      if (v < V) {
        Vec8 w, o;
        w.u = *reinterpret_cast<const uint4*>(gamma + v * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (__half2float(a[i].h[j]) - mean) * rstd;
          const float dxh = __half2float(d[i].h[j]) * __half2float(w.h[j]);
          o.h[j] = __float2half_rn(rstd * (dxh - sa - xh * sb));
        }
        *reinterpret_cast<uint4*>(outr + v * 8) = o.u;
      }
    }
  }
}

// GEGLU backward: din[t, j] = dout * gelu(gate), din[t, I + j] = dout * h * gelu'(gate)
__global__ void __launch_bounds__(256) geglu_bwd_kernel(const __half* __restrict__ in, const __half* __restrict__ dout,
                                                        __half* __restrict__ din, int64_t T, int I) {
  const int VI = I / 8;
  const int64_t nvec = T * VI;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = i / VI;
    const int v = (int)(i % VI);
    const __half* row = in + t * 2 * I;
    Vec8 hh, gg, dd, o1, o2;
    hh.u = *reinterpret_cast<const uint4*>(row + v * 8);
    gg.u = *reinterpret_cast<const uint4*>(row + I + v * 8);
    dd.u = *reinterpret_cast<const uint4*>(dout + i * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float g = __half2float(gg.h[j]), d = __half2float(dd.h[j]), h = __half2float(hh.h[j]);
      const float cdf = 0.5f * (1.f + erff(g * 0.70710678118654752440f));
      const float pdf = 0.39894228040143267794f * __expf(-0.5f * g * g);
      o1.h[j] = __float2half_rn(d * round_half(g * cdf));
      o2.h[j] = __float2half_rn(round_half(d * h) * (cdf + g * pdf));
    }
    __half* orow = din + t * 2 * I;
    *reinterpret_cast<uint4*>(orow + v * 8) = o1.u;
    *reinterpret_cast<uint4*>(orow + I + v * 8) = o2.u;
  }
}

static int gn_splits(int N, int HW) {
  int S = (148 * 4 + N - 1) / N;
  if (S > HW / 64) S = HW / 64;
  if (S < 1) S = 1;
  if (S > 64) S = 64;
  return S;
}

}  // namespace mc

extern "C" int mc_groupnorm_nhwc_stats(const void* workspace, void* stats, int N, int HW, int G, float eps, void* stream) {
  using namespace mc;
  if (!workspace || !stats || N <= 0) {
    set_error("groupnorm_nhwc_stats: null pointer");
    return MC_E_INVALID;
  }
  const int total = N * G;
  groupnorm_finalize_stats_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      (const float*)workspace, (float*)stats, G, gn_splits(N, HW), eps, total);
  count_launch();
  return check_launch("groupnorm_finalize_stats");
}

extern "C" int mc_groupnorm_nhwc_bwd(const void* x, const void* chan_bias, int frames_per_bias_row, const void* dz,
                                     void* dx, const void* stats, const void* gamma, const void* beta, void* workspace,
                                     int64_t workspace_bytes, int N, int HW, int C, int G, int fuse_silu, void* stream) {
  using namespace mc;
  if (!x || !dz || !dx || !stats || !gamma || !beta || !workspace || N <= 0 || HW <= 0) {
    set_error("groupnorm_nhwc_bwd: null pointer or non-positive dims");
    return MC_E_INVALID;
  }
  if (C % 8 != 0 || C % G != 0 || C / 8 > kGnMaxThreads || G > 256) {
    set_error("groupnorm_nhwc_bwd: need C %% 8 == 0, C %% G == 0, C <= 4096, G <= 256 (got C=%d G=%d)", C, G);
    return MC_E_UNSUPPORTED;
  }
  const int S = gn_splits(N, HW);
  if (workspace_bytes < (int64_t)N * S * G * 2 * (int64_t)sizeof(float)) {
    set_error("groupnorm_nhwc_bwd: workspace too small");
    return MC_E_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int threads1 = (C / 8 <= 256) ? 256 : 512;
  const int smem1 = 2 * threads1 * 8 * sizeof(float);
  const __half *xp = (const __half*)x, *cbp = (const __half*)chan_bias, *dzp = (const __half*)dz;
  const __half *gp = (const __half*)gamma, *bp = (const __half*)beta;
  if (fuse_silu)
    groupnorm_bwd_reduce_kernel<true><<<dim3(N, S), threads1, smem1, st>>>(xp, cbp, frames_per_bias_row, dzp,
                                                                          (const float*)stats, gp, bp, (float*)workspace,
                                                                          HW, C, G, S);
  else
    groupnorm_bwd_reduce_kernel<false><<<dim3(N, S), threads1, smem1, st>>>(xp, cbp, frames_per_bias_row, dzp,
                                                                           (const float*)stats, gp, bp, (float*)workspace,
                                                                           HW, C, G, S);
  count_launch();
  int rc = check_launch("groupnorm_bwd_reduce");
  if (rc != MC_OK) return rc;
  int chunks = (148 * 8 + N - 1) / N;
  const int64_t nvec = (int64_t)HW * (C / 8);
  if (chunks > nvec / 256) chunks = (int)(nvec / 256);
  if (chunks < 1) chunks = 1;
  const int smem2 = 4 * G * sizeof(float);
  if (fuse_silu)
    groupnorm_bwd_apply_kernel<true><<<dim3(N, chunks), 256, smem2, st>>>(xp, cbp, frames_per_bias_row, dzp, (__half*)dx,
                                                                         (const float*)stats, (const float*)workspace, gp,
                                                                         bp, HW, C, G, S, chunks);
  else
    groupnorm_bwd_apply_kernel<false><<<dim3(N, chunks), 256, smem2, st>>>(xp, cbp, frames_per_bias_row, dzp, (__half*)dx,
                                                                          (const float*)stats, (const float*)workspace, gp,
                                                                          bp, HW, C, G, S, chunks);
  count_launch();
  return check_launch("groupnorm_bwd_apply");
}

extern "C" int mc_layernorm_bwd(const void* x, const void* dy, void* dx, const void* gamma, int64_t rows, int C, float eps,
                                void* stream) {
  using namespace mc;
  if (!x || !dy || !dx || !gamma || rows <= 0) {
    set_error("layernorm_bwd: null pointer or rows <= 0");
    return MC_E_INVALID;
  }
  if (C % 8 != 0 || C > 1280) {
    set_error("layernorm_bwd: need C %% 8 == 0 and C <= 1280 (got %d)", C);
    return MC_E_UNSUPPORTED;
  }
  int64_t blocks = (rows + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cudaStream_t st = (cudaStream_t)stream;
  const int vpl = (C / 8 + 31) / 32;
#define MC_LNB(V)                                                                                                     \
  layernorm_bwd_kernel<V><<<(unsigned)blocks, 256, 0, st>>>((const __half*)x, (const __half*)dy, (__half*)dx,         \
                                                            (const __half*)gamma, rows, C, eps)
  switch (vpl) {
    case 1: MC_LNB(1); break;
    case 2: MC_LNB(2); break;
    case 3: MC_LNB(3); break;
    case 4: MC_LNB(4); break;
    default: MC_LNB(5); break;
  }
#undef MC_LNB
  count_launch();
  return check_launch("layernorm_bwd");
}

extern "C" int mc_geglu_bwd(const void* in, const void* dout, void* din, int64_t T, int I, void* stream) {
  using namespace mc;
  if (!in || !dout || !din || T <= 0 || I <= 0) {
    set_error("geglu_bwd: null pointer or non-positive dims");
    return MC_E_INVALID;
  }
  if (I % 8 != 0) {
    set_error("geglu_bwd: inner dim must be a multiple of 8 (got %d)", I);
    return MC_E_UNSUPPORTED;
  }
  const int64_t nvec = T * (I / 8);
  int64_t blocks = (nvec + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  geglu_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const __half*)in, (const __half*)dout, (__half*)din,
                                                                        T, I);
  count_launch();
  return check_launch("geglu_bwd");
}
