// Text cross-attention forward on 5th-gen tensor cores (tcgen05) with TMEM accumulators, sm_100a.
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

#include "mc_common.cuh"

namespace mc {

constexpr int kXM = 128;    // query rows per CTA (UMMA M)
constexpr int kXN = 80;     // padded key count (UMMA N of S, K of PV); 77 text tokens
constexpr int kXThreads = 128;

__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // cute::UMMA::SmemDescriptor: start[0,14) lbo[16,30) sbo[32,46) version[46,48)=1 layout_type[61,64)=0 (no swizzle)
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__device__ __forceinline__ uint32_t umma_instr_desc_f16(int M, int N, bool b_mn_major) {
  // cute::UMMA::InstrDescriptor: c_format[4,6)=1 (F32), a/b_format = 0 (F16), a_major[15], b_major[16], n>>3 [17,23), m>>4 [24,29)
  uint32_t d = 0;
  d |= 1u << 4;
  d |= (b_mn_major ? 1u : 0u) << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

struct XAParams {
  const __half *q, *k, *v;
  __half* o;
  int64_t q_sb, q_sr, kv_sb, kv_sr, o_sb, o_sr;  // batch / row strides in elements (head h at column h*DH)
  int B, Nq, Nk, H;
  float scale;
};

template <int DH>
__global__ void __launch_bounds__(kXThreads, 1) cross_attn_fwd_tc_kernel(const XAParams prm) {
  constexpr int DHP = (DH + 15) / 16 * 16;  // head dim padded to the MMA K / N granularity
  constexpr int KCQ = DHP / 8;              // 16-byte chunks per Q / K row
  constexpr int KS1 = DHP / 16;             // k16 steps of S = Q K^T
  constexpr int KS2 = kXN / 16;             // k16 steps of O = P V  (80 / 16 = 5)
  constexpr int TCOLS = 256;                // TMEM columns: S at [0, 80), O at [96, 96 + DHP)
  constexpr int O_COL = 96;
  static_assert(O_COL + DHP <= TCOLS, "TMEM budget");

  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);          // MMA-done barrier
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 16);
  uint8_t* sQ = smem + 128;                                   // K-major [KCQ][128][16 B]
  uint8_t* sK = sQ + KCQ * kXM * 16;                          // K-major [KCQ][80][16 B]
  uint8_t* sV = sK + KCQ * kXN * 16;                          // MN-major [KCQ][80][16 B] (n-chunk c, key j)
  uint8_t* sP = sV + KCQ * kXN * 16;                          // K-major [10][128][16 B]

  const int tid = threadIdx.x, warp = tid >> 5;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * kXM;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }

  // ---- stage Q (row = this thread's query), K and V (rows < 80 by the first 80 threads) as 16-byte chunks ----
  {
    const int row = q0 + tid;
    const __half* qrow = prm.q + (int64_t)b * prm.q_sb + (int64_t)row * prm.q_sr + h * DH;
#pragma unroll
    for (int c = 0; c < KCQ; ++c) {
      uint4 val = make_uint4(0u, 0u, 0u, 0u);
      if (row < prm.Nq && c * 8 < DH) val = *reinterpret_cast<const uint4*>(qrow + c * 8);
      *reinterpret_cast<uint4*>(sQ + (c * kXM + tid) * 16) = val;
    }
    if (tid < kXN) {
      const __half* krow = prm.k + (int64_t)b * prm.kv_sb + (int64_t)tid * prm.kv_sr + h * DH;
      const __half* vrow = prm.v + (int64_t)b * prm.kv_sb + (int64_t)tid * prm.kv_sr + h * DH;
#pragma unroll
      for (int c = 0; c < KCQ; ++c) {
        uint4 kk = make_uint4(0u, 0u, 0u, 0u), vv = make_uint4(0u, 0u, 0u, 0u);
        if (tid < prm.Nk && c * 8 < DH) {
          kk = *reinterpret_cast<const uint4*>(krow + c * 8);
          vv = *reinterpret_cast<const uint4*>(vrow + c * 8);
        }
        *reinterpret_cast<uint4*>(sK + (c * kXN + tid) * 16) = kk;
        *reinterpret_cast<uint4*>(sV + (c * kXN + tid) * 16) = vv;
      }
    }
  }
  fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor-core (async) proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- S = Q K^T ----
  if (tid == 0) {
    const uint32_t idesc = umma_instr_desc_f16(kXM, kXN, false);
#pragma unroll
    for (int ks = 0; ks < KS1; ++ks) {
      const uint64_t a = umma_smem_desc(smem_u32(sQ) + ks * 2 * kXM * 16, kXM * 16, 128);
      const uint64_t bd = umma_smem_desc(smem_u32(sK) + ks * 2 * kXN * 16, kXN * 16, 128);
      umma_f16(tmem_base, a, bd, idesc, ks > 0 ? 1u : 0u);
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();

  // ---- softmax over the 77 valid keys of row `tid` (TMEM lane tid), P -> fp16 -> sP (K-major chunks) ----
  {
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    float s[kXN];
#pragma unroll
    for (int c = 0; c < kXN / 16; ++c) {
      uint32_t r[16];
      tmem_ld16(lane_addr + c * 16, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) s[c * 16 + j] = __uint_as_float(r[j]) * prm.scale;
    }
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < kXN; ++j)
      if (j < prm.Nk) mx = fmaxf(mx, s[j]);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < kXN; ++j) {
      s[j] = (j < prm.Nk) ? __expf(s[j] - mx) : 0.f;
      sum += s[j];
    }
    const float inv = 1.f / sum;
#pragma unroll
    for (int c = 0; c < kXN / 8; ++c) {
      uint4 pk;
      pk.x = pack_half2(s[c * 8 + 0] * inv, s[c * 8 + 1] * inv);
      pk.y = pack_half2(s[c * 8 + 2] * inv, s[c * 8 + 3] * inv);
      pk.z = pack_half2(s[c * 8 + 4] * inv, s[c * 8 + 5] * inv);
      pk.w = pack_half2(s[c * 8 + 6] * inv, s[c * 8 + 7] * inv);
      *reinterpret_cast<uint4*>(sP + (c * kXM + tid) * 16) = pk;
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // ---- O = P V  (A = P K-major; B = V MN-major: n-chunks SBO apart, 8-key groups LBO = 128 B apart) ----
  if (tid == 0) {
    const uint32_t idesc = umma_instr_desc_f16(kXM, DHP, true);
#pragma unroll
    for (int ks = 0; ks < KS2; ++ks) {
      const uint64_t a = umma_smem_desc(smem_u32(sP) + ks * 2 * kXM * 16, kXM * 16, 128);
      const uint64_t bd = umma_smem_desc(smem_u32(sV) + ks * 2 * 128, 128, kXN * 16);
      umma_f16(tmem_base + O_COL, a, bd, idesc, ks > 0 ? 1u : 0u);
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 1);
  tc_fence_after();

  // ---- epilogue: O row `tid` from TMEM -> fp16 -> global ----
  {
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16) + O_COL;
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TCOLS) : "memory");
  }
}

template <int DH>
static int launch_xattn(const XAParams& prm, cudaStream_t st) {
  constexpr int DHP = (DH + 15) / 16 * 16;
  constexpr int KCQ = DHP / 8;
  const int smem = 128 + KCQ * kXM * 16 + 2 * KCQ * kXN * 16 + (kXN / 8) * kXM * 16;
  auto kern = cross_attn_fwd_tc_kernel<DH>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  dim3 grid((prm.Nq + kXM - 1) / kXM, prm.H, prm.B);
  kern<<<grid, kXThreads, smem, st>>>(prm);
  count_launch();
  return check_launch("cross_attn_fwd_tc");
}

}  // namespace mc

extern "C" int mc_cross_attn_fwd(const void* q, const void* k, const void* v, void* o, int B, int Nq, int Nk, int H, int DH,
                                 int64_t q_stride_b, int64_t q_stride_row, int64_t kv_stride_b, int64_t kv_stride_row,
                                 int64_t o_stride_b, int64_t o_stride_row, float scale, void* stream) {
  using namespace mc;
  if (!q || !k || !v || !o || B <= 0 || Nq <= 0 || Nk <= 0 || H <= 0) {
    set_error("cross_attn_fwd: null pointer or non-positive dims");
    return MC_E_INVALID;
  }
  if (Nk > kXN) {
    set_error("cross_attn_fwd: at most %d keys (text tokens) per tile, got %d", kXN, Nk);
    return MC_E_UNSUPPORTED;
  }
  if ((q_stride_row | kv_stride_row | o_stride_row | q_stride_b | kv_stride_b | o_stride_b) % 8 != 0) {
    set_error("cross_attn_fwd: strides must be multiples of 8 elements (16-byte rows)");
    return MC_E_INVALID;
  }
  XAParams prm{(const __half*)q, (const __half*)k, (const __half*)v, (__half*)o, q_stride_b, q_stride_row, kv_stride_b,
               kv_stride_row, o_stride_b, o_stride_row, B, Nq, Nk, H, scale};
  cudaStream_t st = (cudaStream_t)stream;
  switch (DH) {
    case 40: return launch_xattn<40>(prm, st);
    case 80: return launch_xattn<80>(prm, st);
    case 160: return launch_xattn<160>(prm, st);
    case 16: return launch_xattn<16>(prm, st);
    case 32: return launch_xattn<32>(prm, st);
    case 64: return launch_xattn<64>(prm, st);
    default: break;
  }
  set_error("cross_attn_fwd: unsupported head dim %d (16, 32, 40, 64, 80, 160)", DH);
  return MC_E_UNSUPPORTED;
}
