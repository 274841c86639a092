// rb_head_tc.cu -- layer 1 of the fused noisy dueling head on the 5th-generation tensor cores (sm_100a: TMA + mbarrier
// pipeline + tcgen05.mma with the accumulator in tensor memory).
//
//   h[m][s*H + n] = relu( sum_k x[m][k] * W_s[n][k] + b_s[n] ),   W = mu + sigma * (eps_out (outer) eps_in)    (model.py:39-44,73-74)
//
// for the value (s = 0) and advantage (s = 1) streams, M <= 64 rows per tile (the learner's [s; s'] batch), K = 3136.
// This is the one genuinely GEMM-shaped kernel of the update whose operand traffic matters (25.7 MB of mu / sigma per
// pass): FFMA issue limits it to ~0.15 of the HBM roofline (k_head_fc<.,1> in rb_head.cu, kept for shapes this kernel does
// not cover), so it runs on the tensor cores -- in fp32-equivalent arithmetic:
//
//   * orientation: D[128 weight rows x NB batch rows] += A[128 x 8] * B[NB x 8]^T, A = a 128-row slab of W (UMMA M = 128),
//     B = the activations (UMMA N = NB = 32 or 64), accumulator D in TMEM (NB fp32 columns x 128 lanes);
//   * error-compensated TF32 ("3xTF32"): every fp32 operand v is split exactly into hi = v with the low 13 mantissa bits
//     cleared (a TF32 number) and lo = v - hi; D += Whi*Xhi + Wlo*Xhi + Whi*Xlo.  The dropped Wlo*Xlo term is 2^-22 relative,
//     the accumulation is fp32 in TMEM: results agree with an fp32 FMA chain to ~1e-6 relative (tests: <= 2e-5 of scale
//     against cuBLAS fp32, <= 1e-5 on the loss against the reference);
//   * raw mu / sigma / x tiles arrive by TMA (cp.async.bulk.tensor.2d, 128-byte swizzle -> the canonical K-major UMMA
//     layout) through a 4-stage mbarrier ring; eight "compose" warps turn each landed stage IN PLACE into (Whi, Wlo, Xhi,
//     Xlo): W = fma(sigma, eps_out[n]*eps_in[k], mu) exactly as the FFMA kernels compose it, then the split; one elected
//     thread issues the 12 tcgen05.mma of the stage and commits the stage back to the TMA producer;
//   * split-K over ~17 CTAs per 128-row slab (136 CTAs, one per SM); partial tiles go to `part` and k_head_reduce1 sums
//     them in fixed order with bias + ReLU (deterministic; a last-CTA reduction would serialise 0.5 MB per slab on one SM).
//
// Warp roles (320 threads): warp 0 = TMA producer (one lane), warp 1 = TMEM allocator + MMA issuer (one lane),
// warps 2-9 = compose; warps 2-5 then run the epilogue (tcgen05.ld of their 32-lane quarter of the accumulator).

#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "rainbow_b200.h"
#include "rb_internal.cuh"

namespace {

constexpr int TC_BM = 128;        // weight rows per CTA = UMMA M
constexpr int TC_BK = 32;         // k per stage: 32 floats = 128 B = one swizzle-atom row
constexpr int TC_STAGES = 4;
constexpr int TC_COMPOSE = 256;    // compose threads (8 warps)
constexpr int TC_THREADS = 64 + TC_COMPOSE;
constexpr int TC_MAX_PRE = 8;      // k tiles whose eps_in chunk a compose thread keeps in registers
constexpr int TC_W_BYTES = TC_BM * TC_BK * 4;   // 16 KB per weight tile

struct TcArgs {
  const float* eo[2]; const float* ei[2];        // noise factor vectors (null: eval mode)
  const float* bmu[2]; const float* bsg[2];
  float* part;      // [S][M][2H] split-K partials (S > 1)
  float* out;       // [M][2H] (S == 1: written directly with bias + ReLU)
  int K1, H, M, m_lo, kt_total, kt_per, noisy;
};

// ---- PTX wrappers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}\n" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int c0, int c1) {   // pull a box into L2 only
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major operand tile, 128-byte swizzle, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp):
// start address >> 4 in bits [0,14), LBO (ignored for swizzled K-major, canonical value 1) in [16,30), SBO = 1024 >> 4 in
// [32,46), descriptor version 1 (Blackwell) in [46,48), layout type SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (bits [4,6) = 1), A = B = TF32 (bits [7,10), [10,13) = 2),
// both K-major (bits 15, 16 = 0), N >> 3 in [17,23), M >> 4 in [24,29)
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {   // arrives on `bar` when every MMA issued so far has completed
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
      "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// exact split of an fp32 value into a TF32 number (low 13 mantissa bits cleared) and the remainder
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
  lo = __fsub_rn(v, hi);
}
__device__ __forceinline__ void split4(const float4 w, float4& hi, float4& lo) {
  split_tf32(w.x, hi.x, lo.x); split_tf32(w.y, hi.y, lo.y); split_tf32(w.z, hi.z, lo.z); split_tf32(w.w, hi.w, lo.w);
}

template <int NB>   // batch rows per tile = UMMA N (32 or 64)
__global__ void __launch_bounds__(TC_THREADS, 1)
k_head_fc1_tc(const __grid_constant__ CUtensorMap tm_mu0, const __grid_constant__ CUtensorMap tm_mu1,
              const __grid_constant__ CUtensorMap tm_sg0, const __grid_constant__ CUtensorMap tm_sg1,
              const __grid_constant__ CUtensorMap tm_xlo, const __grid_constant__ CUtensorMap tm_xhi,
              const __grid_constant__ TcArgs a) {
  constexpr int X_BYTES = NB * TC_BK * 4;                    // one activation tile (hi or lo)
  constexpr int STAGE_BYTES = 2 * TC_W_BYTES + 2 * X_BYTES;  // Wmu->Whi | Wsigma->Wlo | X->Xhi | Xlo
  constexpr int TMEM_COLS = NB;                              // power of two >= 32
  extern __shared__ uint8_t tc_smem_raw[];
  // 1024-byte alignment (swizzle atoms) by pointer arithmetic on the shared array, so every access stays LDS / STS
  uint8_t* smem = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_STAGES * STAGE_BYTES);
  uint64_t* full = bars;                    // TMA bytes of the stage have landed
  uint64_t* ready = bars + TC_STAGES;       // the stage has been composed / split (128 arrivals)
  uint64_t* empty = bars + 2 * TC_STAGES;   // the stage's MMAs have completed
  uint64_t* acc_full = bars + 3 * TC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * TC_STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_stream = (a.H + TC_BM - 1) / TC_BM;
  const int s = (int)blockIdx.x / tiles_per_stream;
  const int n0 = ((int)blockIdx.x % tiles_per_stream) * TC_BM;
  const int kt_begin = (int)blockIdx.y * a.kt_per;
  const int nkt = min(a.kt_total, kt_begin + a.kt_per) - kt_begin;
  const int m0 = (int)blockIdx.z * NB;
  const bool noisy = a.noisy != 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < TC_STAGES; ++i) {
      mbar_init(full + i, 1);
      mbar_init(ready + i, TC_COMPOSE);
      mbar_init(empty + i, 1);
    }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // TMEM: NB fp32 columns x 128 lanes for the accumulator
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0 && nkt > 0) {
      const CUtensorMap* mu_map = s ? &tm_mu1 : &tm_mu0;
      const CUtensorMap* sg_map = s ? &tm_sg1 : &tm_sg0;
      tma_prefetch_desc(mu_map);
      if (noisy) tma_prefetch_desc(sg_map);
      tma_prefetch_desc(&tm_xlo);
      const uint32_t bytes = TC_W_BYTES * (noisy ? 2 : 1) + X_BYTES;
      // the weight tiles beyond the ring's depth start their trip from HBM to L2 now, so that when a stage frees up its
      // reload is an L2 hit instead of a second exposed DRAM latency
      for (int i = TC_STAGES; i < nkt; ++i) {
        tma_prefetch_l2_2d(mu_map, (kt_begin + i) * TC_BK, n0);
        if (noisy) tma_prefetch_l2_2d(sg_map, (kt_begin + i) * TC_BK, n0);
      }
      for (int i = 0; i < nkt; ++i) {
        const int st = i % TC_STAGES, round = i / TC_STAGES;
        if (i >= TC_STAGES) mbar_wait(empty + st, (round - 1) & 1);
        uint8_t* base = smem + st * STAGE_BYTES;
        mbar_expect_tx(full + st, bytes);
        const int k0 = (kt_begin + i) * TC_BK;
        tma_load_2d(base, mu_map, full + st, k0, n0);
        if (noisy) tma_load_2d(base + TC_W_BYTES, sg_map, full + st, k0, n0);
#pragma unroll
        for (int j = 0; j < NB / 8; ++j) {   // 8-row boxes: each is one 1024-byte swizzle atom; rows past the tensor are zero-filled
          const int r = m0 + 8 * j;
          if (r < a.m_lo) tma_load_2d(base + 2 * TC_W_BYTES + j * 1024, &tm_xlo, full + st, k0, r);
          else tma_load_2d(base + 2 * TC_W_BYTES + j * 1024, &tm_xhi, full + st, k0, r - a.m_lo);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0 && nkt > 0) {
      constexpr uint32_t idesc = umma_idesc(TC_BM, NB);
      for (int i = 0; i < nkt; ++i) {
        const int st = i % TC_STAGES, round = i / TC_STAGES;
        mbar_wait(ready + st, round & 1);
        tc_fence_after();
        const uint32_t base = smem_u32(smem + st * STAGE_BYTES);
        const uint64_t w_hi = umma_desc(base), w_lo = umma_desc(base + TC_W_BYTES);
        const uint64_t x_hi = umma_desc(base + 2 * TC_W_BYTES), x_lo = umma_desc(base + 2 * TC_W_BYTES + X_BYTES);
#pragma unroll
        for (int kk = 0; kk < TC_BK / 8; ++kk) {   // UMMA K = 8 TF32 = 32 bytes: advance the start address inside the swizzle atom
          const uint64_t off = (uint64_t)(kk * 32 >> 4);
          umma_tf32(tmem_base, w_lo + off, x_hi + off, idesc, (i > 0 || kk > 0) ? 1u : 0u);   // small terms first
          umma_tf32(tmem_base, w_hi + off, x_lo + off, idesc, 1u);
          umma_tf32(tmem_base, w_hi + off, x_hi + off, idesc, 1u);
        }
        umma_commit(empty + st);      // implies tcgen05.fence::before_thread_sync
      }
      umma_commit(acc_full);
    }
    __syncwarp();
  } else {
    // ===== compose warps: raw (mu, sigma, x) -> (Whi, Wlo, Xhi, Xlo) in place; then the epilogue =====
    const int ct = threadIdx.x - 64;                     // 0..255
    const int crow = ct >> 3, cphys = ct & 7;            // 16-byte chunk `cphys` of rows crow + 32 j
    const int clog = cphys ^ (crow & 7);                 // its logical position in the row (128-byte swizzle), same for every j
    constexpr int WJ = TC_BM * 8 / TC_COMPOSE;           // float4 of the weight tile per thread (4)
    constexpr int XJ = (NB * 8 + TC_COMPOSE - 1) / TC_COMPOSE;
    const float* __restrict__ eo = a.eo[s];
    const float* __restrict__ ei = a.ei[s];
    float eo_r[WJ];
#pragma unroll
    for (int j = 0; j < WJ; ++j) {
      const int n = n0 + crow + 32 * j;
      eo_r[j] = (noisy && n < a.H) ? __ldg(eo + n) : 0.0f;
    }
    // this thread's eps_in chunk of every k tile of the slice, fetched up front (one memory latency instead of one per stage)
    float4 e_pre[TC_MAX_PRE];
#pragma unroll
    for (int i = 0; i < TC_MAX_PRE; ++i) {
      e_pre[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (noisy && i < nkt) e_pre[i] = __ldg(reinterpret_cast<const float4*>(ei + (size_t)(kt_begin + i) * TC_BK + 4 * clog));
    }
    for (int i = 0; i < nkt; ++i) {
      const int st = i % TC_STAGES, round = i / TC_STAGES;
      float4 e4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < TC_MAX_PRE) {
#pragma unroll
        for (int u = 0; u < TC_MAX_PRE; ++u)
          if (u == i) e4 = e_pre[u];
      } else if (noisy) {
        e4 = __ldg(reinterpret_cast<const float4*>(ei + (size_t)(kt_begin + i) * TC_BK + 4 * clog));
      }
      mbar_wait(full + st, round & 1);
      float4* Wm = reinterpret_cast<float4*>(smem + st * STAGE_BYTES);
      float4* Ws = reinterpret_cast<float4*>(smem + st * STAGE_BYTES + TC_W_BYTES);
      float4* Xa = reinterpret_cast<float4*>(smem + st * STAGE_BYTES + 2 * TC_W_BYTES);
      float4* Xb = reinterpret_cast<float4*>(smem + st * STAGE_BYTES + 2 * TC_W_BYTES + X_BYTES);
#pragma unroll
      for (int j = 0; j < WJ; ++j) {
        const int idx = ct + TC_COMPOSE * j;             // float4 index = row * 8 + physical chunk
        float4 w = Wm[idx];
        if (noisy) {                                     // W = mu + sigma * (eps_out[n] * eps_in[k])   (model.py:39,43)
          const float4 sg = Ws[idx];
          const float e = eo_r[j];
          w.x = fmaf(sg.x, e * e4.x, w.x); w.y = fmaf(sg.y, e * e4.y, w.y);
          w.z = fmaf(sg.z, e * e4.z, w.z); w.w = fmaf(sg.w, e * e4.w, w.w);
        }
        float4 hi, lo;
        split4(w, hi, lo);
        Wm[idx] = hi;
        Ws[idx] = lo;
      }
#pragma unroll
      for (int j = 0; j < XJ; ++j) {
        const int idx = ct + TC_COMPOSE * j;
        if (idx < NB * 8) {
          float4 hi, lo;
          split4(Xa[idx], hi, lo);
          Xa[idx] = hi;
          Xb[idx] = lo;
        }
      }
      fence_proxy_async();          // generic-proxy writes -> visible to the tensor core's (async proxy) reads
      mbar_arrive(ready + st);
    }
    // ----- epilogue: this warp's 32-lane quarter of the accumulator (lane = weight row, column = batch row) -----
    if (warp < 6) {
    if (nkt > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
    }
    const int q = warp & 3;
    const int n = n0 + q * 32 + lane;
    const int ncols = 2 * a.H;
    const int S = gridDim.y;
    float bias = 0.0f;
    if (S == 1 && n < a.H) {
      bias = __ldg(a.bmu[s] + n);
      if (noisy) bias = fmaf(__ldg(a.bsg[s] + n), __ldg(eo + n), bias);
    }
#pragma unroll
    for (int c0 = 0; c0 < NB; c0 += 32) {
      uint32_t v[32];
      if (nkt > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int m = 0; m < 32; ++m) v[m] = 0u;
      }
      if (n < a.H) {
#pragma unroll
        for (int m = 0; m < 32; ++m) {
          const int row = m0 + c0 + m;
          if (row < a.M) {
            const float acc = __uint_as_float(v[m]);
            if (S > 1) __stcg(a.part + ((size_t)blockIdx.y * a.M + row) * ncols + s * a.H + n, acc);
            else a.out[(size_t)row * ncols + s * a.H + n] = fmaxf(acc + bias, 0.0f);
          }
        }
      }
    }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// h[m][c] = relu(sum_s part[s][m][c] + b[c]) in fixed slice order; 4 columns per thread, all slice loads of a batch in flight
__global__ void __launch_bounds__(128)
k_head_reduce1(const float* __restrict__ part, int S, int M, int H, const float* __restrict__ bmu0, const float* __restrict__ bmu1,
               const float* __restrict__ bsg0, const float* __restrict__ bsg1, const float* __restrict__ eo0,
               const float* __restrict__ eo1, float* __restrict__ out) {
  const int ncols = 2 * H;
  const int per_row = ncols >> 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * per_row) return;
  const int m = idx / per_row, c = (idx - m * per_row) << 2;
  const size_t slice = (size_t)M * ncols;
  const float4* src = reinterpret_cast<const float4*>(part + (size_t)m * ncols + c);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s0 = 0; s0 < S; s0 += 24) {             // every slice of the (usual) 17 in flight at once: one round trip
    float4 pv[24];
#pragma unroll
    for (int u = 0; u < 24; ++u)
      pv[u] = (s0 + u < S) ? __ldcg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + (size_t)(s0 + u) * slice))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 24; ++u) { acc.x += pv[u].x; acc.y += pv[u].y; acc.z += pv[u].z; acc.w += pv[u].w; }
  }
  const int st = c >= H ? 1 : 0, n = c - st * H;     // H % 4 == 0: a float4 never straddles the two streams
  const float* bmu = st ? bmu1 : bmu0;
  const float* bsg = st ? bsg1 : bsg0;
  const float* eo = st ? eo1 : eo0;
  float4 b = __ldg(reinterpret_cast<const float4*>(bmu + n));
  if (eo) {
    const float4 bs = __ldg(reinterpret_cast<const float4*>(bsg + n)), e = __ldg(reinterpret_cast<const float4*>(eo + n));
    b.x = fmaf(bs.x, e.x, b.x); b.y = fmaf(bs.y, e.y, b.y); b.z = fmaf(bs.z, e.z, b.z); b.w = fmaf(bs.w, e.w, b.w);
  }
  float4 r;
  r.x = fmaxf(acc.x + b.x, 0.f); r.y = fmaxf(acc.y + b.y, 0.f); r.z = fmaxf(acc.z + b.z, 0.f); r.w = fmaxf(acc.w + b.w, 0.f);
  *reinterpret_cast<float4*>(out + (size_t)m * ncols + c) = r;
}

// ---- host side -------------------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    (void)cudaGetLastError();
  }
  return fn;
}

// 2-D fp32 tensor [rows][cols] (row stride = cols), box = [box_rows][32 columns], 128-byte swizzle, zero fill out of bounds
bool make_map(CUtensorMap* map, const float* base, int rows, int cols, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

namespace rbi {

// Layer-1 split: S slices of kt_per k-tiles (32 columns each) per 128-row slab, aiming at one CTA per SM.
void head_fc1_tc_splits(int K1, int H, int* S, int* kt_per) {
  const int kt = (K1 + TC_BK - 1) / TC_BK;
  const int slabs = 2 * ((H + TC_BM - 1) / TC_BM);
  int want = 148 / slabs;
  if (want < 1) want = 1;
  if (want > kt) want = kt;
  const int per = (kt + want - 1) / want;
  *kt_per = per;
  *S = (kt + per - 1) / per;
}

bool head_fc1_tc_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("RB_HEAD_TC");
    on = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  return on == 1;
}

// Shapes the tensor-core kernel covers: 16-byte aligned rows of whole 32-column k tiles, row blocks of x_lo / x_hi that do
// not straddle an 8-row TMA box, and a driver that exports cuTensorMapEncodeTiled.
bool head_fc1_tc_ok(int K1, int H, int m_lo, int m_hi) {
  if (!head_fc1_tc_enabled()) return false;
  if (K1 % TC_BK || H % 4 || m_lo <= 0) return false;
  if (m_hi > 0 && (m_lo % 8)) return false;
  return encode_fn() != nullptr;
}

int head_fc1_tc(const float* const* w_mu, const float* const* w_sig, const float* const* b_mu, const float* const* b_sig,
                const float* const* ei, const float* const* eo, int K1, int H, const float* x_lo, int m_lo, const float* x_hi,
                int m_hi, float* part, float* h, cudaStream_t st) {
  const int M = m_lo + m_hi;
  const bool noisy = ei[0] != nullptr;
  CUtensorMap tmu[2], tsg[2], txl, txh;
  bool ok = true;
  for (int s = 0; s < 2; ++s) {
    ok = ok && make_map(&tmu[s], w_mu[s], H, K1, TC_BM);
    ok = ok && make_map(&tsg[s], w_sig[s], H, K1, TC_BM);
  }
  ok = ok && make_map(&txl, x_lo, m_lo, K1, 8);
  ok = ok && make_map(&txh, m_hi > 0 ? x_hi : x_lo, m_hi > 0 ? m_hi : m_lo, K1, 8);
  if (!ok) return fail(RB_ERR_CUDA, "rb_head_forward: cuTensorMapEncodeTiled failed");
  TcArgs a;
  for (int s = 0; s < 2; ++s) {
    a.eo[s] = eo[s]; a.ei[s] = ei[s]; a.bmu[s] = b_mu[s]; a.bsg[s] = b_sig[s];
  }
  int S, per;
  head_fc1_tc_splits(K1, H, &S, &per);
  a.part = part; a.out = h; a.K1 = K1; a.H = H; a.M = M; a.m_lo = m_lo;
  a.kt_total = (K1 + TC_BK - 1) / TC_BK; a.kt_per = per; a.noisy = noisy ? 1 : 0;
  const int NB = (M > 32) ? 64 : 32;
  const int slabs = 2 * ((H + TC_BM - 1) / TC_BM);
  dim3 grid(slabs, S, (M + NB - 1) / NB);
  const size_t smem64 = (size_t)TC_STAGES * (2 * TC_W_BYTES + 2 * 64 * TC_BK * 4) + 1024 + 256;
  const size_t smem32 = (size_t)TC_STAGES * (2 * TC_W_BYTES + 2 * 32 * TC_BK * 4) + 1024 + 256;
  int rc = ensure_dynamic_smem(k_head_fc1_tc<64>, smem64, "rb_head_forward(tc)");
  if (rc == RB_OK) rc = ensure_dynamic_smem(k_head_fc1_tc<32>, smem32, "rb_head_forward(tc)");
  if (rc != RB_OK) return rc;
  {
    ProfScope prof_(RB_K_HEAD_FC1, st);
    if (NB == 64) k_head_fc1_tc<64><<<grid, TC_THREADS, smem64, st>>>(tmu[0], tmu[1], tsg[0], tsg[1], txl, txh, a);
    else k_head_fc1_tc<32><<<grid, TC_THREADS, smem32, st>>>(tmu[0], tmu[1], tsg[0], tsg[1], txl, txh, a);
  }
  rc = check_launch("rb_head_forward(fc1 tcgen05)");
  if (rc != RB_OK || S == 1) return rc;
  {
    ProfScope prof_(RB_K_HEAD_REDUCE1, st);
    const int threads = M * (2 * H / 4);
    k_head_reduce1<<<(threads + 127) / 128, 128, 0, st>>>(part, S, M, H, b_mu[0], b_mu[1], b_sig[0], b_sig[1], eo[0], eo[1], h);
  }
  return check_launch("rb_head_forward(reduce1)");
}

}  // namespace rbi
