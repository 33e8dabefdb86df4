// Text cross-attention BACKWARD (dQ) on 5th-gen tensor cores (tcgen05) with TMEM accumulators, sm_100a. The forward is
// csrc/cross_attn_fwd_tc.cu; the shared pieces (softmax_row_tmem, XACfg) below serve the backward's recompute.
//
// Replaces the xformers seam for `attn2` (reference models/attention.py:193-201, :280-285 -> :535-542,
// xformers.ops.memory_efficient_attention): O = softmax(scale * Q K^T) V with Q [b, f*N, C] (all frames of one prompt:
// the text K/V [b, 77, C] are shared by every frame) and 8 heads of DH in {40, 80, 160}.
//
// One CTA = one (batch, head, 128-query tile); the whole key axis (77 -> 80) is ONE tile, so there is no online-softmax
// loop:   S[128 x 80] = Q K^T   (tcgen05.mma, M=128, N=80, K=DH in steps of 16, fp32 accumulators in TMEM)
//         P = softmax(scale * S) row-wise: thread r owns TMEM lane r (tcgen05.ld 32x32b), P -> fp16 -> shared memory
//         O[128 x DH] = P V      (tcgen05.mma, A = P K-major, B = V MN-major straight from its [key][DH] rows)
// Operands sit in shared memory in the no-swizzle "interleave" canonical layouts (8-row x 16-byte core matrices):
//   K-major  tile [R rows][K]: 16-byte chunk (row r, k-chunk c) at  c * (R*16) + r*16   (LBO = R*16, SBO = 128)
//   MN-major tile [K rows][N]: 16-byte chunk (k-row j, n-chunk c) at c * (Kpad*16) + j*16 (SBO = Kpad*16, LBO = 128)
// DH = 40 is padded to 48 with a zero chunk (K of QK^T and N of PV must be multiples of 16).
#include <math.h>

#include "tc_common.cuh"

namespace mc {

constexpr int kXM = 128;    // query rows per CTA (UMMA M)
constexpr int kXN = 80;     // padded key count (UMMA N of S, K of PV); 77 text tokens
constexpr int kXThreads = 128;

struct XAParams {
  const __half *q, *k, *v;
  const __half* d_o;  // bwd only
  __half* o;          // fwd: O; bwd: dQ
  int64_t q_sb, q_sr, kv_sb, kv_sr, o_sb, o_sr, do_sb, do_sr;  // batch / row strides in elements (head h at column h*DH)
  int B, Nq, Nk, H;
  float scale;
};

// softmax(scale * S) of TMEM lane `lane_addr` over the first nk of 80 columns; result as 40 packed half2 (fp16-rounded
// probabilities, zeros beyond nk). Same arithmetic in the forward and in the backward's recompute.
__device__ __forceinline__ void softmax_row_tmem(uint32_t lane_addr, int nk, float scale, uint32_t (&ph)[kXN / 2]) {
  float s[kXN];
#pragma unroll
  for (int c = 0; c < kXN / 16; ++c) {
    uint32_t r[16];
    tmem_ld16(lane_addr + c * 16, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) s[c * 16 + j] = __uint_as_float(r[j]) * scale;
  }
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < kXN; ++j)
    if (j < nk) mx = fmaxf(mx, s[j]);
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < kXN; ++j) {
    s[j] = (j < nk) ? __expf(s[j] - mx) : 0.f;
    sum += s[j];
  }
  const float inv = 1.f / sum;
#pragma unroll
  for (int j = 0; j < kXN / 2; ++j) ph[j] = pack_half2(s[2 * j] * inv, s[2 * j + 1] * inv);
}

template <int DH>
struct XACfg {
  static constexpr int DHP = (DH + 15) / 16 * 16;  // head dim padded to the MMA K / N granularity
  static constexpr int KCQ = DHP / 8;              // 16-byte chunks per Q / K / V / dO row
  static constexpr int KS1 = DHP / 16;             // k16 steps over the head dim
  static constexpr int KS2 = kXN / 16;             // k16 steps over the keys (80 / 16 = 5)
  static constexpr int Q_BYTES = KCQ * kXM * 16;
  static constexpr int KV_BYTES = KCQ * kXN * 16;
  static constexpr int P_BYTES = (kXN / 8) * kXM * 16;
  // forward: O [0, DHP) re-uses the columns of S [0, 80) once every thread holds its probabilities in registers
  static constexpr int TCOLS_FWD = (DHP <= 128) ? 128 : 256;
  static constexpr int SMEM_FWD = 128 + Q_BYTES + 2 * KV_BYTES + P_BYTES;
  // backward: S [0, 80), dP [96, 176); dQ [0, DHP) re-uses both once dS is in shared memory
  static constexpr int DP_COL = 96;
  static constexpr int TCOLS_BWD = 256;
  static constexpr int SMEM_BWD = 128 + 2 * Q_BYTES + 2 * KV_BYTES + P_BYTES;
};

// ================================================================================================================
// backward with respect to Q only: the text K / V come from frozen projections of a constant prompt embedding
// (reference t2v_video_sample.py:67-68; utils/motionclone_functions.py:236 differentiates w.r.t. the latents), so
// dK and dV are never needed on this path.
//   S = Q K^T, dP = dO V^T            (two tcgen05.mma chains, one commit)
//   P = softmax(scale S)  (recomputed, same arithmetic as the forward), D = sum_j P_j dP_j
//   dS = scale * P o (dP - D) -> fp16 -> shared memory
//   dQ = dS K                          (A = dS K-major, B = K MN-major: the SAME shared-memory tile as the K-major
//                                       operand of Q K^T, read through a different descriptor)
// ================================================================================================================
template <int DH>
__global__ void __launch_bounds__(kXThreads) cross_attn_bwd_dq_tc_kernel(const XAParams prm) {
  using X = XACfg<DH>;
  constexpr int DHP = X::DHP, KCQ = X::KCQ;

  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 16);
  uint8_t* sQ = smem + 128;            // K-major [KCQ][128][16 B]
  uint8_t* sD = sQ + X::Q_BYTES;       // dO, K-major [KCQ][128][16 B]
  uint8_t* sK = sD + X::Q_BYTES;       // [KCQ][80][16 B]: K-major for Q K^T, MN-major for dS K
  uint8_t* sV = sK + X::KV_BYTES;      // [KCQ][80][16 B]: K-major for dO V^T
  uint8_t* sS = sV + X::KV_BYTES;      // dS, K-major [10][128][16 B]

  const int tid = threadIdx.x, warp = tid >> 5;
  const int h = blockIdx.x, qt = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * kXM;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(X::TCOLS_BWD)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }

  stage_chunks<DH, KCQ, kXM, kXM, kXThreads>(sQ, prm.q + (int64_t)b * prm.q_sb + h * DH, prm.q_sr, q0, prm.Nq, tid);
  stage_chunks<DH, KCQ, kXM, kXM, kXThreads>(sD, prm.d_o + (int64_t)b * prm.do_sb + h * DH, prm.do_sr, q0, prm.Nq, tid);
  stage_chunks<DH, KCQ, kXN, kXN, kXThreads>(sK, prm.k + (int64_t)b * prm.kv_sb + h * DH, prm.kv_sr, 0, prm.Nk, tid);
  stage_chunks<DH, KCQ, kXN, kXN, kXThreads>(sV, prm.v + (int64_t)b * prm.kv_sb + h * DH, prm.kv_sr, 0, prm.Nk, tid);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- S = Q K^T -> [0, 80);  dP = dO V^T -> [96, 176) ----
  if (tid == 0) {
    const uint32_t idesc = umma_instr_desc_f16(kXM, kXN, false);
#pragma unroll
    for (int ks = 0; ks < X::KS1; ++ks) {
      const uint64_t a = umma_smem_desc(smem_u32(sQ) + ks * 2 * kXM * 16, kXM * 16, 128);
      const uint64_t bd = umma_smem_desc(smem_u32(sK) + ks * 2 * kXN * 16, kXN * 16, 128);
      umma_f16(tmem_base, a, bd, idesc, ks > 0 ? 1u : 0u);
    }
#pragma unroll
    for (int ks = 0; ks < X::KS1; ++ks) {
      const uint64_t a = umma_smem_desc(smem_u32(sD) + ks * 2 * kXM * 16, kXM * 16, 128);
      const uint64_t bd = umma_smem_desc(smem_u32(sV) + ks * 2 * kXN * 16, kXN * 16, 128);
      umma_f16(tmem_base + X::DP_COL, a, bd, idesc, ks > 0 ? 1u : 0u);
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();

  // ---- row `tid`: P (recomputed), D = sum_j P_j dP_j, dS = scale * P (dP - D) -> fp16 -> sS ----
  const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
  {
    uint32_t ph[kXN / 2];
    softmax_row_tmem(lane_addr, prm.Nk, prm.scale, ph);
    float dsum = 0.f;
#pragma unroll
    for (int c = 0; c < kXN / 16; ++c) {
      uint32_t r[16];
      tmem_ld16(lane_addr + X::DP_COL + c * 16, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 p2 = __half22float2(*reinterpret_cast<const __half2*>(&ph[c * 8 + j]));
        dsum = fmaf(p2.x, __uint_as_float(r[2 * j]), dsum);
        dsum = fmaf(p2.y, __uint_as_float(r[2 * j + 1]), dsum);
      }
    }
#pragma unroll
    for (int c = 0; c < kXN / 16; ++c) {
      uint32_t r[16];
      tmem_ld16(lane_addr + X::DP_COL + c * 16, r);
      tmem_ld_wait();
      uint32_t ds[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 p2 = __half22float2(*reinterpret_cast<const __half2*>(&ph[c * 8 + j]));
        ds[j] = pack_half2(prm.scale * p2.x * (__uint_as_float(r[2 * j]) - dsum),
                           prm.scale * p2.y * (__uint_as_float(r[2 * j + 1]) - dsum));
      }
      *reinterpret_cast<uint4*>(sS + ((2 * c) * kXM + tid) * 16) = make_uint4(ds[0], ds[1], ds[2], ds[3]);
      *reinterpret_cast<uint4*>(sS + ((2 * c + 1) * kXM + tid) * 16) = make_uint4(ds[4], ds[5], ds[6], ds[7]);
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();  // all reads of S and dP are complete: dQ may overwrite their columns
  tc_fence_after();

  // ---- dQ = dS K ----
  if (tid == 0) {
    const uint32_t idesc = umma_instr_desc_f16(kXM, DHP, true);
#pragma unroll
    for (int ks = 0; ks < X::KS2; ++ks) {
      const uint64_t a = umma_smem_desc(smem_u32(sS) + ks * 2 * kXM * 16, kXM * 16, 128);
      const uint64_t bd = umma_smem_desc(smem_u32(sK) + ks * 2 * 128, 128, kXN * 16);
      umma_f16(tmem_base, a, bd, idesc, ks > 0 ? 1u : 0u);
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 1);
  tc_fence_after();

  {
    const int row = q0 + tid;
    __half* orow = prm.o + (int64_t)b * prm.o_sb + (int64_t)row * prm.o_sr + h * DH;
#pragma unroll
    for (int c = 0; c < DHP / 16; ++c) {
      uint32_t r[16];
      tmem_ld16(lane_addr + c * 16, r);
      tmem_ld_wait();
      if (row < prm.Nq) {
#pragma unroll
        for (int half8 = 0; half8 < 2; ++half8) {
          if (c * 16 + half8 * 8 < DH) {
            uint4 pk;
            pk.x = pack_half2(__uint_as_float(r[half8 * 8 + 0]), __uint_as_float(r[half8 * 8 + 1]));
            pk.y = pack_half2(__uint_as_float(r[half8 * 8 + 2]), __uint_as_float(r[half8 * 8 + 3]));
            pk.z = pack_half2(__uint_as_float(r[half8 * 8 + 4]), __uint_as_float(r[half8 * 8 + 5]));
            pk.w = pack_half2(__uint_as_float(r[half8 * 8 + 6]), __uint_as_float(r[half8 * 8 + 7]));
            *reinterpret_cast<uint4*>(orow + c * 16 + half8 * 8) = pk;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(X::TCOLS_BWD) : "memory");
  }
}

template <int DH>
static int launch_xattn_bwd(const XAParams& prm, cudaStream_t st) {
  const int smem = XACfg<DH>::SMEM_BWD;
  auto kern = cross_attn_bwd_dq_tc_kernel<DH>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  dim3 grid(prm.H, (prm.Nq + kXM - 1) / kXM, prm.B);
  kern<<<grid, kXThreads, smem, st>>>(prm);
  count_launch();
  return check_launch("cross_attn_bwd_dq_tc");
}

static int xattn_check(const char* what, const void* q, const void* k, const void* v, const void* o, int B, int Nq, int Nk,
                       int H, int64_t strides_or) {
  if (!q || !k || !v || !o || B <= 0 || Nq <= 0 || Nk <= 0 || H <= 0) {
    set_error("%s: null pointer or non-positive dims", what);
    return MC_E_INVALID;
  }
  if (Nk > kXN) {
    set_error("%s: at most %d keys (text tokens) per tile, got %d", what, kXN, Nk);
    return MC_E_UNSUPPORTED;
  }
  if (B > 65535 || (Nq + kXM - 1) / kXM > 65535) {
    set_error("%s: grid too large (B <= 65535, Nq <= 65535 * 128)", what);
    return MC_E_UNSUPPORTED;
  }
  if (strides_or % 8 != 0) {
    set_error("%s: strides must be multiples of 8 elements (16-byte rows)", what);
    return MC_E_INVALID;
  }
  return MC_OK;
}

}  // namespace mc

#define MC_XATTN_DISPATCH(FN)                       \
  switch (DH) {                                     \
    case 8: return FN<8>(prm, st);                  \
    case 16: return FN<16>(prm, st);                \
    case 32: return FN<32>(prm, st);                \
    case 40: return FN<40>(prm, st);                \
    case 64: return FN<64>(prm, st);                \
    case 80: return FN<80>(prm, st);                \
    case 160: return FN<160>(prm, st);              \
    default: break;                                 \
  }

extern "C" int mc_cross_attn_bwd_dq(const void* q, const void* k, const void* v, const void* d_o, void* dq, int B, int Nq,
                                    int Nk, int H, int DH, int64_t q_stride_b, int64_t q_stride_row, int64_t kv_stride_b,
                                    int64_t kv_stride_row, int64_t do_stride_b, int64_t do_stride_row,
                                    int64_t dq_stride_b, int64_t dq_stride_row, float scale, void* stream) {
  using namespace mc;
  const int rc = xattn_check("cross_attn_bwd_dq", q, k, v, dq, B, Nq, Nk, H,
                             q_stride_row | kv_stride_row | dq_stride_row | q_stride_b | kv_stride_b | dq_stride_b |
                                 do_stride_b | do_stride_row);
  if (rc != MC_OK) return rc;
  if (!d_o) {
    set_error("cross_attn_bwd_dq: null d_o");
    return MC_E_INVALID;
  }
  XAParams prm{};
  prm.q = (const __half*)q, prm.k = (const __half*)k, prm.v = (const __half*)v, prm.d_o = (const __half*)d_o;
  prm.o = (__half*)dq;
  prm.q_sb = q_stride_b, prm.q_sr = q_stride_row, prm.kv_sb = kv_stride_b, prm.kv_sr = kv_stride_row;
  prm.o_sb = dq_stride_b, prm.o_sr = dq_stride_row, prm.do_sb = do_stride_b, prm.do_sr = do_stride_row;
  prm.B = B, prm.Nq = Nq, prm.Nk = Nk, prm.H = H, prm.scale = scale;
  cudaStream_t st = (cudaStream_t)stream;
  MC_XATTN_DISPATCH(launch_xattn_bwd)
  set_error("cross_attn_bwd_dq: unsupported head dim %d (8, 16, 32, 40, 64, 80, 160)", DH);
  return MC_E_UNSUPPORTED;
}
