// Spatial self-attention (S1) on 5th-gen tensor cores: tcgen05.mma with TMEM accumulators, operands staged by tensor-map
// TMA (cp.async.bulk.tensor), sm_100a. Forward (this file, part 1) and backward (part 2: dQ kernel, dK/dV kernel).
//
// Replaces the xformers seam of `attn1` (reference models/attention.py:190-192, :271-278 -> :535-542,
// xformers.ops.memory_efficient_attention(q, k, v, attn_bias=None)): O = softmax(scale Q K^T) V per (frame, head) over the
// N = h*w tokens of one frame; N = 4096 / 1024 / 256 / 64 and DH = 40 / 80 / 160 / 160 at 16 x 512 x 512. fp32 softmax
// statistics, one rounding of the output (xformers / flash semantics - SURVEY.md appendix "Attention numerics").
//
// Forward, one CTA = one (frame, head, 128-query tile); 160 threads = 4 softmax warps (thread r owns query row r = TMEM
// lane r) + 1 producer warp whose lane 0 issues every TMA load and every MMA. Two CTAs share an SM (TMEM 2 x 256 columns),
// so one CTA's exponentials overlap the other's MMAs. Per 128-key tile j:
//   producer:  S = Q K_j^T           tcgen05.mma M=128 N=128 K=DH -> TMEM columns [0,128)          -> commit s_full
//   softmax :  S -> registers (128 fp32 per thread), release S (s_free: the producer may issue S_{j+1} at once),
//              row max / exp2 / row sum, P -> fp16 -> shared memory (K-major SW128)                -> arrive p_full
//   producer:  O += P V_j            tcgen05.mma M=128 N=DH K=128 (V MN-major: no transpose)       -> commit pv_done
// O stays in TMEM for the whole key loop. The running maximum is only raised when a row's maximum grows by more than
// 2^8 (P <= 256 fits fp16; exactness is unaffected because the row sum uses the same reference maximum); only then does
// the softmax warp rescale its 32 rows of O in TMEM (tcgen05.ld -> multiply -> tcgen05.st) before releasing P.
// K and V are single-buffered: K_{j+1} is requested the moment S_j has completed and V_{j+1} when P V_j has - both land
// long before the exp-bound softmax of tile j (>= 1024 cycles: 16 384 exponentials at 16 / cycle / SM) is through.
#include <math.h>

#include "tma_common.cuh"

namespace mc {

constexpr int kFM = 128;         // query rows per CTA (UMMA M)
constexpr int kFN = 128;         // keys per tile (UMMA N of S, K extent of P V)
constexpr int kFThreads = 160;   // 4 softmax warps + 1 producer warp
constexpr float kRescaleThreshold = 8.f;  // log2 units

struct FAParams {
  float* lse;        // [B][H][N] natural-log sum-exp of the scaled scores (nullable)
  __half* o;
  int64_t o_sb, o_sr;
  int B, N, H;
  float scale_log2e;  // scale * log2(e)
};

template <int DH>
struct FACfg {
  using T = TileParts<DH>;
  static constexpr int DHP = T::DHP;
  static constexpr int P_BYTES = 2 * 16384;                 // P [128 q][128 keys] fp16: two K-major SW128 parts
  static constexpr int OFF_Q = 0, OFF_K = T::BYTES, OFF_V = 2 * T::BYTES, OFF_P = 3 * T::BYTES;
  static constexpr int OFF_BAR = OFF_P + P_BYTES;
  static constexpr int SMEM = OFF_BAR + 128 + 1024;          // + alignment slack (dynamic smem base is 16 B aligned)
  static constexpr int O_COL = 128;                          // O at TMEM columns [128, 128 + DHP)
  static constexpr int TCOLS = (128 + DHP <= 256) ? 256 : 512;
  static constexpr int CTAS_PER_SM = (TCOLS == 256 && 2 * SMEM <= 227 * 1024) ? 2 : 1;
};

// S tile: A = Q (K-major), B = K (K-major); one MMA per k16 step over the head dim
template <int DH>
__device__ __forceinline__ void issue_qk(uint32_t d_tmem, uint32_t sA, uint32_t sB, int n_rows_b = kFN) {
  using T = TileParts<DH>;
  const uint32_t idesc = umma_idesc_f16(kFM, n_rows_b, false, false);
  uint32_t acc = 0;
#pragma unroll
  for (int p = 0; p < T::N64; ++p)
#pragma unroll
    for (int ks = 0; ks < T::KS64; ++ks) {
      umma_f16(d_tmem, desc_k128(sA + T::part64_off(p), ks), desc_k128(sB + T::part64_off(p), ks), idesc, acc);
      acc = 1;
    }
#pragma unroll
  for (int p = 0; p < T::N16; ++p) {
    umma_f16(d_tmem, desc_k32(sA + T::part16_off(p)), desc_k32(sB + T::part16_off(p)), idesc, acc);
    acc = 1;
  }
}

// D[128 x DH] (+)= A[128 x 128] B[128 x DH]: A = two K-major SW128 parts written by threads (P or dS, K = 128 rows of B),
// B = an operand tile read MN-major (its rows are the K dimension). One MMA per (k16 step, part of B).
template <int DH>
__device__ __forceinline__ void issue_pv(uint32_t d_tmem, uint32_t sA, uint32_t sB, bool accumulate, int ksteps = kFN / 16) {
  using T = TileParts<DH>;
  const uint32_t idesc64 = umma_idesc_f16(kFM, T::W64, false, true);
  const uint32_t idesc16 = umma_idesc_f16(kFM, 16, false, true);
  for (int ks = 0; ks < ksteps; ++ks) {
    const uint64_t a = desc_k128(sA + (ks >> 2) * 16384, ks & 3);
    const uint32_t acc = (accumulate || ks > 0) ? 1u : 0u;
#pragma unroll
    for (int p = 0; p < T::N64; ++p) umma_f16(d_tmem + p * 64, a, desc_mn128(sB + T::part64_off(p), ks), idesc64, acc);
#pragma unroll
    for (int p = 0; p < T::N16; ++p)
      umma_f16(d_tmem + T::N64 * 64 + p * 16, a, desc_mn32(sB + T::part16_off(p), ks), idesc16, acc);
  }
}

template <int DH>
__global__ void __launch_bounds__(kFThreads, FACfg<DH>::CTAS_PER_SM)
spatial_attn_fwd_kernel(const __grid_constant__ CUtensorMap mq128, const __grid_constant__ CUtensorMap mk128,
                        const __grid_constant__ CUtensorMap mv128, const __grid_constant__ CUtensorMap mq32,
                        const __grid_constant__ CUtensorMap mk32, const __grid_constant__ CUtensorMap mv32,
                        const FAParams prm) {
  using X = FACfg<DH>;
  using T = TileParts<DH>;
  constexpr int DHP = X::DHP;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem + X::OFF_Q;
  uint8_t* sK = smem + X::OFF_K;
  uint8_t* sV = smem + X::OFF_V;
  uint8_t* sP = smem + X::OFF_P;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + X::OFF_BAR);
  uint64_t* bar_q = bars + 0;      // Q landed                       (tx)
  uint64_t* bar_k = bars + 1;      // K_j landed                     (tx, phase j & 1)
  uint64_t* bar_v = bars + 2;      // V_j landed                     (tx)
  uint64_t* s_full = bars + 3;     // S_j in TMEM                    (tcgen05.commit)
  uint64_t* s_free = bars + 4;     // S_j copied to registers        (4 warp arrivals)
  uint64_t* p_full = bars + 5;     // P_j in shared memory           (4 warp arrivals)
  uint64_t* pv_done = bars + 6;    // O += P_j V_j completed         (tcgen05.commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * kFM, N = prm.N;
  const int T_tiles = (N + kFN - 1) / kFN;

  if (warp == 4) {
    tmem_alloc<X::TCOLS>(tmem_slot);
    if (lane == 0) {
      mbar_init(bar_q, 1), mbar_init(bar_k, 1), mbar_init(bar_v, 1), mbar_init(s_full, 1);
      mbar_init(s_free, 4), mbar_init(p_full, 4), mbar_init(pv_done, 1);
      fence_mbar_init();
      tma_prefetch_desc(&mq128), tma_prefetch_desc(&mk128), tma_prefetch_desc(&mv128);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ================= producer: TMA + MMA issue (one thread) =================
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_q, T::BYTES);
      tma_load_tile<DH>(sQ, &mq128, &mq32, bar_q, q0, h, b);
      mbar_arrive_expect_tx(bar_k, T::BYTES);
      tma_load_tile<DH>(sK, &mk128, &mk32, bar_k, 0, h, b);
      mbar_arrive_expect_tx(bar_v, T::BYTES);
      tma_load_tile<DH>(sV, &mv128, &mv32, bar_v, 0, h, b);
      mbar_wait(bar_q, 0);
      mbar_wait(bar_k, 0);
      tc_fence_after();
      issue_qk<DH>(tmem_base, smem_u32(sQ), smem_u32(sK));
      umma_commit(s_full);
      for (int j = 0; j < T_tiles; ++j) {
        const uint32_t ph = j & 1;
        mbar_wait(s_full, ph);  // S_j completed: the K buffer is free
        if (j + 1 < T_tiles) {
          mbar_arrive_expect_tx(bar_k, T::BYTES);
          tma_load_tile<DH>(sK, &mk128, &mk32, bar_k, (j + 1) * kFN, h, b);
          mbar_wait(bar_k, ph ^ 1);
          mbar_wait(s_free, ph);  // every softmax thread holds S_j in registers
          tc_fence_after();
          issue_qk<DH>(tmem_base, smem_u32(sQ), smem_u32(sK));
          umma_commit(s_full);
        }
        mbar_wait(bar_v, ph);
        mbar_wait(p_full, ph);
        tc_fence_after();
        issue_pv<DH>(tmem_base + X::O_COL, smem_u32(sP), smem_u32(sV), j > 0);
        umma_commit(pv_done);
        if (j + 1 < T_tiles) {
          mbar_wait(pv_done, ph);  // V buffer (and P buffer) free
          mbar_arrive_expect_tx(bar_v, T::BYTES);
          tma_load_tile<DH>(sV, &mv128, &mv32, bar_v, (j + 1) * kFN, h, b);
        }
      }
    }
  } else {
    // ================= softmax warps: thread = query row =================
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const float c = prm.scale_log2e;
    float m_used = -INFINITY, l = 0.f;
    for (int j = 0; j < T_tiles; ++j) {
      const uint32_t ph = j & 1;
      mbar_wait(s_full, ph);
      tc_fence_after();
      uint32_t s[kFN];
      tmem_ld32(lane_addr + 0, s + 0);
      tmem_ld32(lane_addr + 32, s + 32);
      tmem_ld32(lane_addr + 64, s + 64);
      tmem_ld32(lane_addr + 96, s + 96);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);

      const int kvalid = N - j * kFN;  // keys of this tile that exist (>= 1)
      if (kvalid < kFN) {
#pragma unroll
        for (int i = 0; i < kFN; ++i)
          if (i >= kvalid) s[i] = 0xff800000u;  // -inf
      }
      float mx0 = __uint_as_float(s[0]), mx1 = __uint_as_float(s[1]), mx2 = __uint_as_float(s[2]),
            mx3 = __uint_as_float(s[3]);
#pragma unroll
      for (int i = 4; i < kFN; i += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(s[i])), mx1 = fmaxf(mx1, __uint_as_float(s[i + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(s[i + 2])), mx3 = fmaxf(mx3, __uint_as_float(s[i + 3]));
      }
      const float mxc = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * c;
      bool waited_pv = false;
      if (j == 0) {
        m_used = mxc;
      } else {
        const bool grow = mxc - m_used > kRescaleThreshold;
        if (__any_sync(0xffffffffu, grow)) {
          const float m_new = grow ? mxc : m_used;
          const float alpha = ex2_approx(m_used - m_new);
          l *= alpha;
          m_used = m_new;
          mbar_wait(pv_done, ph ^ 1);  // O holds tiles 0..j-1
          waited_pv = true;
          tc_fence_after();
#pragma unroll
          for (int cc = 0; cc < DHP / 16; ++cc) {
            uint32_t r[16];
            tmem_ld16(lane_addr + X::O_COL + cc * 16, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            tmem_st16(lane_addr + X::O_COL + cc * 16, r);
          }
          tmem_st_wait();
          tc_fence_before();
        }
      }
      // p = exp2(s*c - m_used); packed to fp16 pairs in place (s[0..63] hold the 128 probabilities)
      const float negm = -m_used;
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int i = 0; i < kFN; i += 2) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(s[i]), c, negm));
        const float p1 = ex2_approx(fmaf(__uint_as_float(s[i + 1]), c, negm));
        l0 += p0, l1 += p1;
        s[i >> 1] = pack_half2(p0, p1);
      }
      l += l0 + l1;
      if (j > 0 && !waited_pv) mbar_wait(pv_done, ph ^ 1);  // P_{j-1} consumed: the P buffer is free
#pragma unroll
      for (int ch = 0; ch < 16; ++ch) {  // 16-byte chunk ch = keys [8 ch, 8 ch + 8)
        uint8_t* dst = sP + (ch >> 3) * 16384 + sw128_chunk_off(tid, ch & 7);
        *reinterpret_cast<uint4*>(dst) = make_uint4(s[4 * ch], s[4 * ch + 1], s[4 * ch + 2], s[4 * ch + 3]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // ---- epilogue: O / l -> fp16 -> global; log-sum-exp for the backward ----
    mbar_wait(pv_done, (T_tiles - 1) & 1);
    tc_fence_after();
    const int row = q0 + tid;
    const float inv = 1.f / l;
    __half* orow = prm.o + (int64_t)b * prm.o_sb + (int64_t)row * prm.o_sr + h * DH;
#pragma unroll
    for (int cc = 0; cc < DHP / 16; ++cc) {
      uint32_t r[16];
      tmem_ld16(lane_addr + X::O_COL + cc * 16, r);
      tmem_ld_wait();
      if (row < N) {
#pragma unroll
        for (int half8 = 0; half8 < 2; ++half8) {
          if (cc * 16 + half8 * 8 < DH) {
            uint4 pk;
            pk.x = pack_half2(__uint_as_float(r[half8 * 8 + 0]) * inv, __uint_as_float(r[half8 * 8 + 1]) * inv);
            pk.y = pack_half2(__uint_as_float(r[half8 * 8 + 2]) * inv, __uint_as_float(r[half8 * 8 + 3]) * inv);
            pk.z = pack_half2(__uint_as_float(r[half8 * 8 + 4]) * inv, __uint_as_float(r[half8 * 8 + 5]) * inv);
            pk.w = pack_half2(__uint_as_float(r[half8 * 8 + 6]) * inv, __uint_as_float(r[half8 * 8 + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + cc * 16 + half8 * 8) = pk;
          }
        }
      }
    }
    if (prm.lse != nullptr && row < N)
      prm.lse[((int64_t)b * prm.H + h) * N + row] = (m_used + log2f(l)) * 0.6931471805599453f;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<X::TCOLS>(tmem_base);
}

struct AttnMaps {
  CUtensorMap m128, m32;
};

// maps for one operand tensor; the SW32 map is only encoded when the head dim has 16-wide parts
template <int DH>
static int make_maps(AttnMaps& m, const void* base, int H, int N, int B, int64_t sr, int64_t sb) {
  using T = TileParts<DH>;
  int rc = make_attn_tensor_map(&m.m128, base, DH, H, N, B, sr, sb, 64, 128, true);
  if (rc) return rc;
  if (T::N16 > 0) rc = make_attn_tensor_map(&m.m32, base, DH, H, N, B, sr, sb, 16, 128, false);
  else m.m32 = m.m128;
  return rc;
}

template <int DH>
static int launch_spatial_fwd(const void* q, const void* k, const void* v, const FAParams& prm, int64_t q_sb, int64_t q_sr,
                              int64_t k_sb, int64_t k_sr, int64_t v_sb, int64_t v_sr, cudaStream_t st) {
  using X = FACfg<DH>;
  AttnMaps mq, mk, mv;
  if (make_maps<DH>(mq, q, prm.H, prm.N, prm.B, q_sr, q_sb) || make_maps<DH>(mk, k, prm.H, prm.N, prm.B, k_sr, k_sb) ||
      make_maps<DH>(mv, v, prm.H, prm.N, prm.B, v_sr, v_sb)) {
    set_error("spatial_attn_fwd: cuTensorMapEncodeTiled failed (pointers must be 16-byte aligned, strides multiples of 8)");
    return MC_E_CUDA;
  }
  auto kern = spatial_attn_fwd_kernel<DH>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, X::SMEM);
  dim3 grid((prm.N + kFM - 1) / kFM, prm.H, prm.B);
  kern<<<grid, kFThreads, X::SMEM, st>>>(mq.m128, mk.m128, mv.m128, mq.m32, mk.m32, mv.m32, prm);
  count_launch();
  return check_launch("spatial_attn_fwd");
}

}  // namespace mc

extern "C" int mc_spatial_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int N, int H,
                                   int DH, int64_t q_stride_b, int64_t q_stride_row, int64_t k_stride_b,
                                   int64_t k_stride_row, int64_t v_stride_b, int64_t v_stride_row, int64_t o_stride_b,
                                   int64_t o_stride_row, float scale, void* stream) {
  using namespace mc;
  if (!q || !k || !v || !o || B <= 0 || N <= 0 || H <= 0) {
    set_error("spatial_attn_fwd: null pointer or non-positive dims");
    return MC_E_INVALID;
  }
  if (B > 65535 || H > 65535) {
    set_error("spatial_attn_fwd: at most 65535 frames / heads");
    return MC_E_UNSUPPORTED;
  }
  if ((q_stride_b | q_stride_row | k_stride_b | k_stride_row | v_stride_b | v_stride_row | o_stride_b | o_stride_row) % 8 ||
      ((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)o) % 16) {
    set_error("spatial_attn_fwd: pointers must be 16-byte aligned and strides multiples of 8 elements");
    return MC_E_INVALID;
  }
  FAParams prm{};
  prm.lse = lse, prm.o = (__half*)o, prm.o_sb = o_stride_b, prm.o_sr = o_stride_row;
  prm.B = B, prm.N = N, prm.H = H;
  prm.scale_log2e = scale * 1.44269504088896340736f;
  cudaStream_t st = (cudaStream_t)stream;
#define MC_SA_CASE(D) \
  case D: return launch_spatial_fwd<D>(q, k, v, prm, q_stride_b, q_stride_row, k_stride_b, k_stride_row, v_stride_b, v_stride_row, st);
  switch (DH) {
    MC_SA_CASE(8) MC_SA_CASE(16) MC_SA_CASE(32) MC_SA_CASE(40) MC_SA_CASE(64) MC_SA_CASE(80) MC_SA_CASE(160)
    default: break;
  }
#undef MC_SA_CASE
  set_error("spatial_attn_fwd: unsupported head dim %d (8, 16, 32, 40, 64, 80, 160)", DH);
  return MC_E_UNSUPPORTED;
}

namespace mc {

// =====================================================================================================================
// Part 2: backward (autograd of the seam above, traversed by torch.autograd.grad at utils/motionclone_functions.py:236).
// With P = softmax(scale S), S = Q K^T, D_r = sum_e dO_re O_re:
//     dV = P^T dO      dP = dO V^T      dS = scale * P o (dP - D)      dQ = dS K      dK = dS^T Q
// P is recomputed from the forward's log-sum-exp (no N x N tensor is ever stored). Two kernels, both with the forward's
// structure (4 compute warps, thread = TMEM lane = tile row; lane 0 of a 5th warp issues every TMA load and MMA; TMEM
// accumulators; 2 CTAs per SM where the TMEM budget allows):
//   dQ kernel  : CTA = 128 queries, loop over 64-key tiles:   S, dP (M=128 q, N=64 keys)  ->  dS -> smem  -> dQ += dS K
//   dKV kernel : CTA = 128 keys,   loop over 64-query tiles:  S^T = K Q^T, dP^T = V dO^T (M=128 keys, N=64 q)
//                -> P^T, dS^T -> smem -> dV += P^T dO, dK += dS^T Q
// Every [rows][DH] tile is used by two GEMMs through two descriptors: K-major where DH is the contraction (S, dP) and
// MN-major where the rows are (dS K, P^T dO, dS^T Q) - no transposes, no second copy. The deterministic two-kernel split
// recomputes S and dP once more than a fused kernel would, but needs no atomics on dQ.
// =====================================================================================================================
constexpr int kBT = 64;  // streamed tile height (keys in the dQ kernel, queries in the dKV kernel)

struct FABwdParams {
  const float* lse;    // [B][H][N] from the forward
  const float* dsum;   // [B][H][N]  D = rowsum(dO o O)
  __half *dq, *dk, *dv;
  int64_t g_sb, g_sr;  // dq / dk / dv share one stride pattern (column blocks of one fused gradient buffer, or separate)
  int B, N, H;
  float scale, scale_log2e;
};

// D[b][h][r] = sum_e dO[b][r][h][e] * O[b][r][h][e]
template <int DH>
__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(const __half* __restrict__ o, const __half* __restrict__ d_o,
                                                            float* __restrict__ dsum, int64_t o_sb, int64_t o_sr,
                                                            int64_t do_sb, int64_t do_sr, int B, int N, int H) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // ((b * N) + r) * H + h
  if (i >= (int64_t)B * N * H) return;
  const int h = (int)(i % H);
  const int64_t br = i / H;
  const int r = (int)(br % N), b = (int)(br / N);
  const uint4* po = reinterpret_cast<const uint4*>(o + b * o_sb + (int64_t)r * o_sr + h * DH);
  const uint4* pd = reinterpret_cast<const uint4*>(d_o + b * do_sb + (int64_t)r * do_sr + h * DH);
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < DH / 8; ++c) {
    const uint4 a = po[c], g = pd[c];
    const __half2* ah = reinterpret_cast<const __half2*>(&a);
    const __half2* gh = reinterpret_cast<const __half2*>(&g);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 x = __half22float2(ah[j]), y = __half22float2(gh[j]);
      acc = fmaf(x.x, y.x, acc);
      acc = fmaf(x.y, y.y, acc);
    }
  }
  dsum[((int64_t)b * H + h) * N + r] = acc;
}

// D[128 x DH] (+)= A[128 x 64] B[64 x DH]: A = one K-major SW128 part written by threads, B = a 64-row tile read MN-major
template <int DH>
__device__ __forceinline__ void issue_ab64(uint32_t d_tmem, uint32_t sA, uint32_t sB, bool accumulate) {
  using T = TileParts<DH, kBT>;
  const uint32_t idesc64 = umma_idesc_f16(128, T::W64, false, true);
  const uint32_t idesc16 = umma_idesc_f16(128, 16, false, true);
#pragma unroll
  for (int ks = 0; ks < kBT / 16; ++ks) {
    const uint64_t a = desc_k128(sA, ks);
    const uint32_t acc = (accumulate || ks > 0) ? 1u : 0u;
#pragma unroll
    for (int p = 0; p < T::N64; ++p) umma_f16(d_tmem + p * 64, a, desc_mn128(sB + T::part64_off(p), ks), idesc64, acc);
#pragma unroll
    for (int p = 0; p < T::N16; ++p)
      umma_f16(d_tmem + T::N64 * 64 + p * 16, a, desc_mn32(sB + T::part16_off(p), ks), idesc16, acc);
  }
}

// D[128 x 64] = A[128 x DH] B[64 x DH]^T: A a 128-row tile, B a 64-row tile, both K-major
template <int DH>
__device__ __forceinline__ void issue_qk64(uint32_t d_tmem, uint32_t sA, uint32_t sB) {
  using TA = TileParts<DH, 128>;
  using TB = TileParts<DH, kBT>;
  const uint32_t idesc = umma_idesc_f16(128, kBT, false, false);
  uint32_t acc = 0;
#pragma unroll
  for (int p = 0; p < TA::N64; ++p)
#pragma unroll
    for (int ks = 0; ks < TA::KS64; ++ks) {
      umma_f16(d_tmem, desc_k128(sA + TA::part64_off(p), ks), desc_k128(sB + TB::part64_off(p), ks), idesc, acc);
      acc = 1;
    }
#pragma unroll
  for (int p = 0; p < TA::N16; ++p) {
    umma_f16(d_tmem, desc_k32(sA + TA::part16_off(p)), desc_k32(sB + TB::part16_off(p)), idesc, acc);
    acc = 1;
  }
}

template <int DH>
struct FABwdCfg {
  using TA = TileParts<DH, 128>;   // resident tiles
  using TB = TileParts<DH, kBT>;   // streamed tiles
  static constexpr int DHP = TA::DHP;
  static constexpr int X_BYTES = 128 * 128;  // one K-major SW128 part [128 rows][64] written by threads
  // dQ kernel: Q, dO resident; K, V double-buffered; dS
  static constexpr int DQ_OFF_Q = 0, DQ_OFF_DO = TA::BYTES, DQ_OFF_K = 2 * TA::BYTES, DQ_OFF_V = DQ_OFF_K + 2 * TB::BYTES;
  static constexpr int DQ_OFF_DS = DQ_OFF_V + 2 * TB::BYTES, DQ_OFF_BAR = DQ_OFF_DS + X_BYTES;
  static constexpr int DQ_SMEM = DQ_OFF_BAR + 128 + 1024;
  static constexpr int DQ_COL = 128;  // S [0,64) dP [64,128) dQ [128, 128 + DHP)
  static constexpr int DQ_TCOLS = (128 + DHP <= 256) ? 256 : 512;
  static constexpr int DQ_CTAS = (DQ_TCOLS == 256 && 2 * DQ_SMEM <= 227 * 1024) ? 2 : 1;
  // dKV kernel: K, V resident; Q, dO double-buffered; P^T, dS^T
  static constexpr int KV_OFF_K = 0, KV_OFF_V = TA::BYTES, KV_OFF_Q = 2 * TA::BYTES, KV_OFF_DO = KV_OFF_Q + 2 * TB::BYTES;
  static constexpr int KV_OFF_PT = KV_OFF_DO + 2 * TB::BYTES, KV_OFF_DST = KV_OFF_PT + X_BYTES;
  static constexpr int KV_OFF_BAR = KV_OFF_DST + X_BYTES, KV_SMEM = KV_OFF_BAR + 128 + 1024;
  static constexpr int DV_COL = 128, DK_COL = 128 + DHP;  // S^T [0,64) dP^T [64,128) dV, dK
  static constexpr int KV_TCOLS = (128 + 2 * DHP <= 256) ? 256 : 512;
  static constexpr int KV_CTAS = (KV_TCOLS == 256 && 2 * KV_SMEM <= 227 * 1024) ? 2 : 1;
};

// TMEM row (DHP fp32 columns) -> fp16 -> global row
template <int DH, int DHP>
__device__ __forceinline__ void store_row_from_tmem(uint32_t taddr, __half* grow, bool valid, float mul) {
#pragma unroll
  for (int cc = 0; cc < DHP / 16; ++cc) {
    uint32_t r[16];
    tmem_ld16(taddr + cc * 16, r);
    tmem_ld_wait();
    if (valid) {
#pragma unroll
      for (int half8 = 0; half8 < 2; ++half8) {
        if (cc * 16 + half8 * 8 < DH) {
          uint4 pk;
          pk.x = pack_half2(__uint_as_float(r[half8 * 8 + 0]) * mul, __uint_as_float(r[half8 * 8 + 1]) * mul);
          pk.y = pack_half2(__uint_as_float(r[half8 * 8 + 2]) * mul, __uint_as_float(r[half8 * 8 + 3]) * mul);
          pk.z = pack_half2(__uint_as_float(r[half8 * 8 + 4]) * mul, __uint_as_float(r[half8 * 8 + 5]) * mul);
          pk.w = pack_half2(__uint_as_float(r[half8 * 8 + 6]) * mul, __uint_as_float(r[half8 * 8 + 7]) * mul);
          *reinterpret_cast<uint4*>(grow + cc * 16 + half8 * 8) = pk;
        }
      }
    }
  }
}

// ------------------------------------------------ dQ ------------------------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(kFThreads, FABwdCfg<DH>::DQ_CTAS)
spatial_attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap mq128, const __grid_constant__ CUtensorMap mq32,
                           const __grid_constant__ CUtensorMap mdo128, const __grid_constant__ CUtensorMap mdo32,
                           const __grid_constant__ CUtensorMap mk128, const __grid_constant__ CUtensorMap mk32,
                           const __grid_constant__ CUtensorMap mv128, const __grid_constant__ CUtensorMap mv32,
                           const FABwdParams prm) {
  using X = FABwdCfg<DH>;
  using TA = typename X::TA;
  using TB = typename X::TB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem + X::DQ_OFF_Q;
  uint8_t* sDO = smem + X::DQ_OFF_DO;
  uint8_t* sK = smem + X::DQ_OFF_K;   // 2 stages
  uint8_t* sV = smem + X::DQ_OFF_V;   // 2 stages
  uint8_t* sDS = smem + X::DQ_OFF_DS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + X::DQ_OFF_BAR);
  uint64_t* bar_q = bars + 0;        // Q and dO landed
  uint64_t* bar_kv = bars + 1;       // [2] K_j, V_j landed in stage j & 1
  uint64_t* sdp_full = bars + 3;     // S_j, dP_j in TMEM
  uint64_t* sdp_free = bars + 4;     // copied to registers (4 warp arrivals)
  uint64_t* ds_full = bars + 5;      // dS_j in shared memory (4 warp arrivals)
  uint64_t* dq_done = bars + 6;      // dQ += dS_j K_j completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * kFM, N = prm.N;
  const int T_tiles = (N + kBT - 1) / kBT;

  if (warp == 4) {
    tmem_alloc<X::DQ_TCOLS>(tmem_slot);
    if (lane == 0) {
      mbar_init(bar_q, 1), mbar_init(bar_kv, 1), mbar_init(bar_kv + 1, 1), mbar_init(sdp_full, 1);
      mbar_init(sdp_free, 4), mbar_init(ds_full, 4), mbar_init(dq_done, 1);
      fence_mbar_init();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_q, 2 * TA::BYTES);
      tma_load_tile<DH, 128>(sQ, &mq128, &mq32, bar_q, q0, h, b);
      tma_load_tile<DH, 128>(sDO, &mdo128, &mdo32, bar_q, q0, h, b);
      for (int j = 0; j < 2 && j < T_tiles; ++j) {
        mbar_arrive_expect_tx(bar_kv + j, 2 * TB::BYTES);
        tma_load_tile<DH, kBT>(sK + j * TB::BYTES, &mk128, &mk32, bar_kv + j, j * kBT, h, b);
        tma_load_tile<DH, kBT>(sV + j * TB::BYTES, &mv128, &mv32, bar_kv + j, j * kBT, h, b);
      }
      mbar_wait(bar_q, 0);
      mbar_wait(bar_kv, 0);
      tc_fence_after();
      issue_qk64<DH>(tmem_base, smem_u32(sQ), smem_u32(sK));
      issue_qk64<DH>(tmem_base + 64, smem_u32(sDO), smem_u32(sV));
      umma_commit(sdp_full);
      for (int j = 0; j < T_tiles; ++j) {
        const uint32_t ph = j & 1, st = j & 1;
        if (j + 1 < T_tiles) {
          const int sn = (j + 1) & 1;
          mbar_wait(bar_kv + sn, ((j + 1) >> 1) & 1);
          mbar_wait(sdp_free, ph);
          tc_fence_after();
          issue_qk64<DH>(tmem_base, smem_u32(sQ), smem_u32(sK + sn * TB::BYTES));
          issue_qk64<DH>(tmem_base + 64, smem_u32(sDO), smem_u32(sV + sn * TB::BYTES));
          umma_commit(sdp_full);
        }
        mbar_wait(ds_full, ph);
        tc_fence_after();
        issue_ab64<DH>(tmem_base + X::DQ_COL, smem_u32(sDS), smem_u32(sK + st * TB::BYTES), j > 0);
        umma_commit(dq_done);
        if (j + 2 < T_tiles) {
          mbar_wait(dq_done, ph);  // K_j / V_j consumed: refill the stage with tile j + 2
          mbar_arrive_expect_tx(bar_kv + st, 2 * TB::BYTES);
          tma_load_tile<DH, kBT>(sK + st * TB::BYTES, &mk128, &mk32, bar_kv + st, (j + 2) * kBT, h, b);
          tma_load_tile<DH, kBT>(sV + st * TB::BYTES, &mv128, &mv32, bar_kv + st, (j + 2) * kBT, h, b);
        }
      }
    }
  } else {
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int row = q0 + tid;
    const bool rvalid = row < N;
    const int64_t srow = ((int64_t)b * prm.H + h) * N + (rvalid ? row : 0);
    const float lse2 = rvalid ? prm.lse[srow] * 1.44269504088896340736f : 0.f;
    const float Dr = rvalid ? prm.dsum[srow] : 0.f;
    const float c = prm.scale_log2e, sc = prm.scale;
    for (int j = 0; j < T_tiles; ++j) {
      const uint32_t ph = j & 1;
      mbar_wait(sdp_full, ph);
      tc_fence_after();
      uint32_t s[kBT], dp[kBT];
      tmem_ld32(lane_addr + 0, s), tmem_ld32(lane_addr + 32, s + 32);
      tmem_ld32(lane_addr + 64, dp), tmem_ld32(lane_addr + 96, dp + 32);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sdp_free);
      const int kvalid = N - j * kBT;
#pragma unroll
      for (int i = 0; i < kBT; i += 2) {
        float p0 = ex2_approx(fmaf(__uint_as_float(s[i]), c, -lse2));
        float p1 = ex2_approx(fmaf(__uint_as_float(s[i + 1]), c, -lse2));
        if (i >= kvalid) p0 = 0.f;
        if (i + 1 >= kvalid) p1 = 0.f;
        const float d0 = p0 * (__uint_as_float(dp[i]) - Dr) * sc;
        const float d1 = p1 * (__uint_as_float(dp[i + 1]) - Dr) * sc;
        s[i >> 1] = pack_half2(d0, d1);
      }
      if (j > 0) mbar_wait(dq_done, ph ^ 1);  // dS_{j-1} consumed
#pragma unroll
      for (int ch = 0; ch < 8; ++ch)
        *reinterpret_cast<uint4*>(sDS + sw128_chunk_off(tid, ch)) = make_uint4(s[4 * ch], s[4 * ch + 1], s[4 * ch + 2], s[4 * ch + 3]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full);
    }
    mbar_wait(dq_done, (T_tiles - 1) & 1);
    tc_fence_after();
    store_row_from_tmem<DH, X::DHP>(lane_addr + X::DQ_COL, prm.dq + (int64_t)b * prm.g_sb + (int64_t)row * prm.g_sr + h * DH,
                                    rvalid, 1.f);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    __syncwarp();
    tmem_dealloc<X::DQ_TCOLS>(tmem_base);
  }
}

// ------------------------------------------------ dK, dV --------------------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(kFThreads, FABwdCfg<DH>::KV_CTAS)
spatial_attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap mk128, const __grid_constant__ CUtensorMap mk32,
                            const __grid_constant__ CUtensorMap mv128, const __grid_constant__ CUtensorMap mv32,
                            const __grid_constant__ CUtensorMap mq128, const __grid_constant__ CUtensorMap mq32,
                            const __grid_constant__ CUtensorMap mdo128, const __grid_constant__ CUtensorMap mdo32,
                            const FABwdParams prm) {
  using X = FABwdCfg<DH>;
  using TA = typename X::TA;
  using TB = typename X::TB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sK = smem + X::KV_OFF_K;
  uint8_t* sV = smem + X::KV_OFF_V;
  uint8_t* sQ = smem + X::KV_OFF_Q;    // 2 stages
  uint8_t* sDO = smem + X::KV_OFF_DO;  // 2 stages
  uint8_t* sPT = smem + X::KV_OFF_PT;
  uint8_t* sDST = smem + X::KV_OFF_DST;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + X::KV_OFF_BAR);
  uint64_t* bar_kv = bars + 0;      // K, V landed
  uint64_t* bar_q = bars + 1;       // [2] Q_i, dO_i landed in stage i & 1
  uint64_t* st_full = bars + 3;     // S^T_i, dP^T_i in TMEM
  uint64_t* st_free = bars + 4;     // copied to registers (4 warp arrivals)
  uint64_t* pt_full = bars + 5;     // P^T_i, dS^T_i in shared memory (4 warp arrivals)
  uint64_t* dkv_done = bars + 6;    // dV, dK updates of tile i completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int k0 = kt * kFM, N = prm.N;
  const int T_tiles = (N + kBT - 1) / kBT;

  if (warp == 4) {
    tmem_alloc<X::KV_TCOLS>(tmem_slot);
    if (lane == 0) {
      mbar_init(bar_kv, 1), mbar_init(bar_q, 1), mbar_init(bar_q + 1, 1), mbar_init(st_full, 1);
      mbar_init(st_free, 4), mbar_init(pt_full, 4), mbar_init(dkv_done, 1);
      fence_mbar_init();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_kv, 2 * TA::BYTES);
      tma_load_tile<DH, 128>(sK, &mk128, &mk32, bar_kv, k0, h, b);
      tma_load_tile<DH, 128>(sV, &mv128, &mv32, bar_kv, k0, h, b);
      for (int i = 0; i < 2 && i < T_tiles; ++i) {
        mbar_arrive_expect_tx(bar_q + i, 2 * TB::BYTES);
        tma_load_tile<DH, kBT>(sQ + i * TB::BYTES, &mq128, &mq32, bar_q + i, i * kBT, h, b);
        tma_load_tile<DH, kBT>(sDO + i * TB::BYTES, &mdo128, &mdo32, bar_q + i, i * kBT, h, b);
      }
      mbar_wait(bar_kv, 0);
      mbar_wait(bar_q, 0);
      tc_fence_after();
      issue_qk64<DH>(tmem_base, smem_u32(sK), smem_u32(sQ));
      issue_qk64<DH>(tmem_base + 64, smem_u32(sV), smem_u32(sDO));
      umma_commit(st_full);
      for (int i = 0; i < T_tiles; ++i) {
        const uint32_t ph = i & 1, st = i & 1;
        if (i + 1 < T_tiles) {
          const int sn = (i + 1) & 1;
          mbar_wait(bar_q + sn, ((i + 1) >> 1) & 1);
          mbar_wait(st_free, ph);
          tc_fence_after();
          issue_qk64<DH>(tmem_base, smem_u32(sK), smem_u32(sQ + sn * TB::BYTES));
          issue_qk64<DH>(tmem_base + 64, smem_u32(sV), smem_u32(sDO + sn * TB::BYTES));
          umma_commit(st_full);
        }
        mbar_wait(pt_full, ph);
        tc_fence_after();
        issue_ab64<DH>(tmem_base + X::DV_COL, smem_u32(sPT), smem_u32(sDO + st * TB::BYTES), i > 0);
        issue_ab64<DH>(tmem_base + X::DK_COL, smem_u32(sDST), smem_u32(sQ + st * TB::BYTES), i > 0);
        umma_commit(dkv_done);
        if (i + 2 < T_tiles) {
          mbar_wait(dkv_done, ph);
          mbar_arrive_expect_tx(bar_q + st, 2 * TB::BYTES);
          tma_load_tile<DH, kBT>(sQ + st * TB::BYTES, &mq128, &mq32, bar_q + st, (i + 2) * kBT, h, b);
          tma_load_tile<DH, kBT>(sDO + st * TB::BYTES, &mdo128, &mdo32, bar_q + st, (i + 2) * kBT, h, b);
        }
      }
    }
  } else {
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int row = k0 + tid;  // key row
    const float c = prm.scale_log2e, sc = prm.scale;
    const float* lse_h = prm.lse + ((int64_t)b * prm.H + h) * N;
    const float* dsum_h = prm.dsum + ((int64_t)b * prm.H + h) * N;
    for (int i = 0; i < T_tiles; ++i) {
      const uint32_t ph = i & 1;
      mbar_wait(st_full, ph);
      tc_fence_after();
      uint32_t s[kBT], dp[kBT];
      tmem_ld32(lane_addr + 0, s), tmem_ld32(lane_addr + 32, s + 32);
      tmem_ld32(lane_addr + 64, dp), tmem_ld32(lane_addr + 96, dp + 32);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(st_free);
      const int q0 = i * kBT;
      const int qvalid = N - q0;  // queries of this tile that exist
#pragma unroll
      for (int qq = 0; qq < kBT; qq += 2) {
        // per-query statistics: the same address for every thread of the warp (broadcast, L1-resident)
        const int i0 = q0 + (qq < qvalid ? qq : 0), i1 = q0 + (qq + 1 < qvalid ? qq + 1 : 0);
        const float l0 = __ldg(lse_h + i0) * 1.44269504088896340736f, l1 = __ldg(lse_h + i1) * 1.44269504088896340736f;
        const float D0 = __ldg(dsum_h + i0), D1 = __ldg(dsum_h + i1);
        float p0 = ex2_approx(fmaf(__uint_as_float(s[qq]), c, -l0));
        float p1 = ex2_approx(fmaf(__uint_as_float(s[qq + 1]), c, -l1));
        if (qq >= qvalid) p0 = 0.f;
        if (qq + 1 >= qvalid) p1 = 0.f;
        const float d0 = p0 * (__uint_as_float(dp[qq]) - D0) * sc;
        const float d1 = p1 * (__uint_as_float(dp[qq + 1]) - D1) * sc;
        s[qq >> 1] = pack_half2(p0, p1);
        dp[qq >> 1] = pack_half2(d0, d1);
      }
      if (i > 0) mbar_wait(dkv_done, ph ^ 1);
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const uint32_t off = sw128_chunk_off(tid, ch);
        *reinterpret_cast<uint4*>(sPT + off) = make_uint4(s[4 * ch], s[4 * ch + 1], s[4 * ch + 2], s[4 * ch + 3]);
        *reinterpret_cast<uint4*>(sDST + off) = make_uint4(dp[4 * ch], dp[4 * ch + 1], dp[4 * ch + 2], dp[4 * ch + 3]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(pt_full);
    }
    mbar_wait(dkv_done, (T_tiles - 1) & 1);
    tc_fence_after();
    const bool rvalid = row < N;
    const int64_t goff = (int64_t)b * prm.g_sb + (int64_t)row * prm.g_sr + h * DH;
    store_row_from_tmem<DH, X::DHP>(lane_addr + X::DV_COL, prm.dv + goff, rvalid, 1.f);
    store_row_from_tmem<DH, X::DHP>(lane_addr + X::DK_COL, prm.dk + goff, rvalid, 1.f);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    __syncwarp();
    tmem_dealloc<X::KV_TCOLS>(tmem_base);
  }
}

template <int DH>
static int make_maps_rows(AttnMaps& m, const void* base, int H, int N, int B, int64_t sr, int64_t sb, int rows) {
  using T = TileParts<DH>;
  int rc = make_attn_tensor_map(&m.m128, base, DH, H, N, B, sr, sb, 64, rows, true);
  if (rc) return rc;
  if (T::N16 > 0) rc = make_attn_tensor_map(&m.m32, base, DH, H, N, B, sr, sb, 16, rows, false);
  else m.m32 = m.m128;
  return rc;
}

template <int DH>
static int launch_spatial_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o, float* dsum,
                              const FABwdParams& prm, int64_t q_sb, int64_t q_sr, int64_t k_sb, int64_t k_sr, int64_t v_sb,
                              int64_t v_sr, int64_t o_sb, int64_t o_sr, int64_t do_sb, int64_t do_sr, cudaStream_t st) {
  using X = FABwdCfg<DH>;
  const int B = prm.B, N = prm.N, H = prm.H;
  AttnMaps q128, do128, k128, v128, q64, do64, k64, v64;
  int rc = make_maps_rows<DH>(q128, q, H, N, B, q_sr, q_sb, 128) | make_maps_rows<DH>(do128, d_o, H, N, B, do_sr, do_sb, 128) |
           make_maps_rows<DH>(k128, k, H, N, B, k_sr, k_sb, 128) | make_maps_rows<DH>(v128, v, H, N, B, v_sr, v_sb, 128) |
           make_maps_rows<DH>(q64, q, H, N, B, q_sr, q_sb, kBT) | make_maps_rows<DH>(do64, d_o, H, N, B, do_sr, do_sb, kBT) |
           make_maps_rows<DH>(k64, k, H, N, B, k_sr, k_sb, kBT) | make_maps_rows<DH>(v64, v, H, N, B, v_sr, v_sb, kBT);
  if (rc) {
    set_error("spatial_attn_bwd: cuTensorMapEncodeTiled failed (pointers must be 16-byte aligned, strides multiples of 8)");
    return MC_E_CUDA;
  }
  {
    const int64_t total = (int64_t)B * N * H;
    attn_bwd_prep_kernel<DH><<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const __half*)o, (const __half*)d_o, dsum, o_sb,
                                                                            o_sr, do_sb, do_sr, B, N, H);
    count_launch();
    if (int e = check_launch("attn_bwd_prep")) return e;
  }
  dim3 grid((N + kFM - 1) / kFM, H, B);
  {
    auto kern = spatial_attn_bwd_dq_kernel<DH>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, X::DQ_SMEM);
    kern<<<grid, kFThreads, X::DQ_SMEM, st>>>(q128.m128, q128.m32, do128.m128, do128.m32, k64.m128, k64.m32, v64.m128, v64.m32, prm);
    count_launch();
    if (int e = check_launch("spatial_attn_bwd_dq")) return e;
  }
  {
    auto kern = spatial_attn_bwd_dkv_kernel<DH>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, X::KV_SMEM);
    kern<<<grid, kFThreads, X::KV_SMEM, st>>>(k128.m128, k128.m32, v128.m128, v128.m32, q64.m128, q64.m32, do64.m128, do64.m32, prm);
    count_launch();
    if (int e = check_launch("spatial_attn_bwd_dkv")) return e;
  }
  return MC_OK;
}

}  // namespace mc

extern "C" int64_t mc_spatial_attn_bwd_workspace_bytes(int B, int N, int H) { return (int64_t)B * N * H * 4; }

extern "C" int mc_spatial_attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                                   const float* lse, void* dq, void* dk, void* dv, void* workspace, int B, int N, int H,
                                   int DH, int64_t q_stride_b, int64_t q_stride_row, int64_t k_stride_b,
                                   int64_t k_stride_row, int64_t v_stride_b, int64_t v_stride_row, int64_t o_stride_b,
                                   int64_t o_stride_row, int64_t do_stride_b, int64_t do_stride_row, int64_t g_stride_b,
                                   int64_t g_stride_row, float scale, void* stream) {
  using namespace mc;
  if (!q || !k || !v || !o || !d_o || !lse || !dq || !dk || !dv || !workspace || B <= 0 || N <= 0 || H <= 0) {
    set_error("spatial_attn_bwd: null pointer or non-positive dims");
    return MC_E_INVALID;
  }
  if (B > 65535 || H > 65535) {
    set_error("spatial_attn_bwd: at most 65535 frames / heads");
    return MC_E_UNSUPPORTED;
  }
  if ((q_stride_b | q_stride_row | k_stride_b | k_stride_row | v_stride_b | v_stride_row | o_stride_b | o_stride_row |
       do_stride_b | do_stride_row | g_stride_b | g_stride_row) % 8 ||
      ((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)o | (uintptr_t)d_o | (uintptr_t)dq | (uintptr_t)dk |
       (uintptr_t)dv) % 16) {
    set_error("spatial_attn_bwd: pointers must be 16-byte aligned and strides multiples of 8 elements");
    return MC_E_INVALID;
  }
  FABwdParams prm{};
  prm.lse = lse, prm.dsum = (const float*)workspace;
  prm.dq = (__half*)dq, prm.dk = (__half*)dk, prm.dv = (__half*)dv, prm.g_sb = g_stride_b, prm.g_sr = g_stride_row;
  prm.B = B, prm.N = N, prm.H = H, prm.scale = scale, prm.scale_log2e = scale * 1.44269504088896340736f;
  cudaStream_t st = (cudaStream_t)stream;
#define MC_SB_CASE(D)                                                                                                       \
  case D:                                                                                                                   \
    return launch_spatial_bwd<D>(q, k, v, o, d_o, (float*)workspace, prm, q_stride_b, q_stride_row, k_stride_b, k_stride_row, \
                                 v_stride_b, v_stride_row, o_stride_b, o_stride_row, do_stride_b, do_stride_row, st);
  switch (DH) {
    MC_SB_CASE(8) MC_SB_CASE(16) MC_SB_CASE(32) MC_SB_CASE(40) MC_SB_CASE(64) MC_SB_CASE(80) MC_SB_CASE(160)
    default: break;
  }
#undef MC_SB_CASE
  set_error("spatial_attn_bwd: unsupported head dim %d (8, 16, 32, 40, 64, 80, 160)", DH);
  return MC_E_UNSUPPORTED;
}
