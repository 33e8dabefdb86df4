// Spatial self-attention for SHORT sequences (N <= 256 tokens per frame) on tcgen05 tensor cores with TMEM accumulators,
// sm_100a.
//
// Replaces the xformers seam for `attn1` (reference models/attention.py:190-192, :271-278 -> :535-542,
// xformers.ops.memory_efficient_attention) at the two deepest UNet levels: 16x16 and 8x8 latent pixels per frame
// (N = 256 / 64), 8 heads of DH = 160. (The library flash kernels have no sm_100 instantiation for DH = 160 and fall back
// to sm_80 wmma code: 170 us per call at N = 256 against a ~10 us roofline; profiles/README.md.)
//
// One CTA = one (frame, head, 128-query tile). The WHOLE key axis (<= 256) is one TMEM tile, so there is no online-softmax
// loop:   S[128 x 256] = Q K^T          tcgen05.mma M=128, N=256, K=DH (fp32 accumulators: 256 TMEM columns)
//         K is dead once S is committed -> V streams into the same shared-memory buffer with cp.async WHILE the
//         softmax runs (thread r owns TMEM lane r: pass 1 row max, pass 2 exp2 / sum, P -> fp16 -> shared memory)
//         O[128 x DH] = P V             tcgen05.mma, A = P K-major, B = V MN-major; O re-uses S's TMEM columns
//         epilogue: O / rowsum -> fp16 -> global
// Operand layouts: the no-swizzle canonical layouts of tc_common.cuh; DH is padded to a multiple of 16 with zero chunks.
#include <math.h>

#include "tc_common.cuh"

namespace mc {

constexpr int kSM = 128;       // query rows per CTA (UMMA M)
constexpr int kSNK = 256;      // padded key count (UMMA N of S, K extent of P V)
constexpr int kSThreads = 128;

struct SAParams {
  const __half *q, *k, *v;
  __half* o;
  int64_t qkv_sb, qkv_sr, o_sb, o_sr;  // frame / token strides in elements (q, k, v share them; head h at column h*DH)
  int B, N, H;
  float scale_log2e;
};

__device__ __forceinline__ void cp_async16_zfill(void* sdst, const void* gsrc, bool valid) {
  const uint32_t sz = valid ? 16u : 0u;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(sdst)), "l"(gsrc), "r"(sz) : "memory");
}

template <int DH>
struct SACfg {
  static constexpr int DHP = (DH + 15) / 16 * 16;
  static constexpr int KC = DH / 8;     // real 16-byte chunks per row
  static constexpr int KCQ = DHP / 8;   // padded
  static constexpr int KS1 = DHP / 16;  // k16 steps over the head dim
  static constexpr int Q_BYTES = KCQ * kSM * 16;
  static constexpr int KV_BYTES = KCQ * kSNK * 16;
  static constexpr int P_BYTES = (kSNK / 8) * kSM * 16;
  static constexpr int SMEM = 128 + Q_BYTES + KV_BYTES + P_BYTES;
  static constexpr int TCOLS = 256;  // S [0, 256); O [0, DHP) once P is in shared memory
};

template <int DH>
__global__ void __launch_bounds__(kSThreads) self_attn_short_tc_kernel(const SAParams prm) {
  using X = SACfg<DH>;
  constexpr int DHP = X::DHP, KC = X::KC, KCQ = X::KCQ;

  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);  // MMA-done barrier
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 16);
  uint8_t* sQ = smem + 128;            // K-major [KCQ][128][16 B]
  uint8_t* sKV = sQ + X::Q_BYTES;      // K: K-major [KCQ][256][16 B]; then V: MN-major, same addressing
  uint8_t* sP = sKV + X::KV_BYTES;     // K-major [32][128][16 B]

  const int tid = threadIdx.x, warp = tid >> 5;
  const int h = blockIdx.x, qt = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * kSM, N = prm.N;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(X::TCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  const int64_t base = (int64_t)b * prm.qkv_sb + h * DH;
  stage_chunks<DH, KCQ, kSM, kSM, kSThreads>(sQ, prm.q + base, prm.qkv_sr, q0, N, tid);
  stage_chunks<DH, KCQ, kSNK, kSNK, kSThreads>(sKV, prm.k + base, prm.qkv_sr, 0, N, tid);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- S = Q K^T ----
  if (tid == 0) {
    const uint32_t idesc = umma_instr_desc_f16(kSM, kSNK, false);
#pragma unroll
    for (int ks = 0; ks < X::KS1; ++ks) {
      const uint64_t a = umma_smem_desc(smem_u32(sQ) + ks * 2 * kSM * 16, kSM * 16, 128);
      const uint64_t bd = umma_smem_desc(smem_u32(sKV) + ks * 2 * kSNK * 16, kSNK * 16, 128);
      umma_f16(tmem_base, a, bd, idesc, ks > 0 ? 1u : 0u);
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();

  // ---- K is dead: V -> sKV (cp.async, zero-filled beyond N; the padding chunks are still zero from K's staging) ----
  {
    const __half* vsrc = prm.v + base;
    for (int i = tid; i < kSNK * KC; i += kSThreads) {
      const int r = i / KC, c = i % KC;
      const bool valid = r < N;
      cp_async16_zfill(sKV + (c * kSNK + r) * 16, vsrc + (valid ? (int64_t)r * prm.qkv_sr + c * 8 : 0), valid);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }

  // ---- softmax of row `tid` (TMEM lane tid) over the N valid keys; unnormalised P -> fp16 -> sP ----
  const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
  float mx = -INFINITY;
#pragma unroll 4
  for (int c = 0; c < kSNK / 16; ++c) {
    if (c * 16 < N) {
      uint32_t r[16];
      tmem_ld16(lane_addr + c * 16, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (c * 16 + j < N) mx = fmaxf(mx, __uint_as_float(r[j]));
    }
  }
  float sum = 0.f;
#pragma unroll 2
  for (int c = 0; c < kSNK / 16; ++c) {
    uint32_t pk[8];
    if (c * 16 < N) {
      uint32_t r[16];
      tmem_ld16(lane_addr + c * 16, r);
      tmem_ld_wait();
      float p[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        p[j] = (c * 16 + j < N) ? exp2f((__uint_as_float(r[j]) - mx) * prm.scale_log2e) : 0.f;
        sum += p[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) pk[j] = pack_half2(p[2 * j], p[2 * j + 1]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) pk[j] = 0u;
    }
    *reinterpret_cast<uint4*>(sP + ((2 * c) * kSM + tid) * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    *reinterpret_cast<uint4*>(sP + ((2 * c + 1) * kSM + tid) * 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();  // P and V are in shared memory; every tcgen05.ld of S has completed (O overwrites its columns)
  tc_fence_after();

  // ---- O = P V over the key k16-steps that hold valid keys ----
  if (tid == 0) {
    const uint32_t idesc = umma_instr_desc_f16(kSM, DHP, true);
    const int nks = (N + 15) / 16;
    for (int ks = 0; ks < nks; ++ks) {
      const uint64_t a = umma_smem_desc(smem_u32(sP) + ks * 2 * kSM * 16, kSM * 16, 128);
      const uint64_t bd = umma_smem_desc(smem_u32(sKV) + ks * 2 * 128, 128, kSNK * 16);
      umma_f16(tmem_base, a, bd, idesc, ks > 0 ? 1u : 0u);
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 1);
  tc_fence_after();

  // ---- epilogue: O row `tid` / rowsum -> fp16 -> global ----
  {
    const float inv = 1.f / sum;
    const int row = q0 + tid;
    __half* orow = prm.o + (int64_t)b * prm.o_sb + (int64_t)row * prm.o_sr + h * DH;
#pragma unroll
    for (int c = 0; c < DHP / 16; ++c) {
      uint32_t r[16];
      tmem_ld16(lane_addr + c * 16, r);
      tmem_ld_wait();
      if (row < N) {
#pragma unroll
        for (int half8 = 0; half8 < 2; ++half8) {
          if (c * 16 + half8 * 8 < DH) {
            uint4 pk;
            pk.x = pack_half2(__uint_as_float(r[half8 * 8 + 0]) * inv, __uint_as_float(r[half8 * 8 + 1]) * inv);
            pk.y = pack_half2(__uint_as_float(r[half8 * 8 + 2]) * inv, __uint_as_float(r[half8 * 8 + 3]) * inv);
            pk.z = pack_half2(__uint_as_float(r[half8 * 8 + 4]) * inv, __uint_as_float(r[half8 * 8 + 5]) * inv);
            pk.w = pack_half2(__uint_as_float(r[half8 * 8 + 6]) * inv, __uint_as_float(r[half8 * 8 + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + c * 16 + half8 * 8) = pk;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(X::TCOLS) : "memory");
  }
}

template <int DH>
static int launch_self_short(const SAParams& prm, cudaStream_t st) {
  const int smem = SACfg<DH>::SMEM;
  auto kern = self_attn_short_tc_kernel<DH>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  dim3 grid(prm.H, (prm.N + kSM - 1) / kSM, prm.B);
  kern<<<grid, kSThreads, smem, st>>>(prm);
  count_launch();
  return check_launch("self_attn_short_tc");
}

}  // namespace mc

extern "C" int mc_self_attn_short_fwd(const void* q, const void* k, const void* v, void* o, int B, int N, int H, int DH,
                                      int64_t qkv_stride_b, int64_t qkv_stride_row, int64_t o_stride_b,
                                      int64_t o_stride_row, float scale, void* stream) {
  using namespace mc;
  if (!q || !k || !v || !o || B <= 0 || N <= 0 || H <= 0) {
    set_error("self_attn_short_fwd: null pointer or non-positive dims");
    return MC_E_INVALID;
  }
  if (N > kSNK || B > 65535) {
    set_error("self_attn_short_fwd: at most %d tokens per frame and 65535 frames (got N=%d B=%d)", kSNK, N, B);
    return MC_E_UNSUPPORTED;
  }
  if ((qkv_stride_b | qkv_stride_row | o_stride_b | o_stride_row) % 8 != 0) {
    set_error("self_attn_short_fwd: strides must be multiples of 8 elements (16-byte rows)");
    return MC_E_INVALID;
  }
  SAParams prm{};
  prm.q = (const __half*)q, prm.k = (const __half*)k, prm.v = (const __half*)v, prm.o = (__half*)o;
  prm.qkv_sb = qkv_stride_b, prm.qkv_sr = qkv_stride_row, prm.o_sb = o_stride_b, prm.o_sr = o_stride_row;
  prm.B = B, prm.N = N, prm.H = H;
  prm.scale_log2e = scale * 1.44269504088896340736f;
  cudaStream_t st = (cudaStream_t)stream;
  switch (DH) {
    case 40: return launch_self_short<40>(prm, st);
    case 64: return launch_self_short<64>(prm, st);
    case 80: return launch_self_short<80>(prm, st);
    case 160: return launch_self_short<160>(prm, st);
    default: break;
  }
  set_error("self_attn_short_fwd: unsupported head dim %d (40, 64, 80, 160)", DH);
  return MC_E_UNSUPPORTED;
}
