// rb_peer.cu -- fused gradient exchange + optimiser over NVLink peer memory (multi-GPU learners).
//
// Replaces  ncclAllReduce(flat_grad) ; k_sqnorm ; k_clip_adam (replicated on every rank)  by
//   k_peer_reduce : reduce-scatter done with peer LOADS -- rank r sums slice r of every rank's gradient buffer
//                   (fixed rank order), keeps the sum locally and accumulates its squared norm;
//   k_peer_adam   : clip coefficient from the W per-rank partial norms, Adam on the owned slice only (the moments
//                   are sharded: 1/W of the optimiser traffic per GPU), all-gather done with peer STORES -- the
//                   updated parameter slice is written straight into every rank's flat parameter buffer;
//   k_peer_fence  : waits until every rank's slice has landed (and therefore every rank is done reading this
//                   rank's gradients), so plain stream order protects the next forward / backward.
// Cross-GPU ordering uses monotonically increasing epoch flags in peer-visible memory (st.release.sys /
// ld.acquire.sys); nothing spins on the host.  All buffers handed in as `peer_*[rank]` pointers must be mapped on
// every GPU (CUDA IPC or torch symmetric memory).
//
// ROUND-1 STATUS: compiled only.  Not exercised on hardware (the round's GPU budget was spent before a 2-GPU
// validation slot was left); rainbow_b200 does not call it unless Agent(..., peer_optimizer=True).

#include <cuda_runtime.h>
#include <stdint.h>

#include "rainbow_b200.h"
#include "rb_internal.cuh"

namespace {

constexpr int PEER_THREADS = 256;
constexpr int PEER_MAX_CTAS = 148 * 4;

struct PeerBufs {
  const float* grad[RB_MAX_PEERS];   // every rank's flat gradient buffer (read)
  float* param[RB_MAX_PEERS];        // every rank's flat parameter buffer (written)
  uint64_t* flags[RB_MAX_PEERS];     // every rank's flag block: [0..W) grads ready, [W..2W) norm ready, [2W..3W) params written
  double* norms[RB_MAX_PEERS];       // every rank's norm block: [W] partial squared norms
  int world, rank;
};

__device__ __forceinline__ void st_release_sys(uint64_t* p, uint64_t v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// one thread per CTA waits until the W local flags of `block` reach `epoch`, then the CTA proceeds
__device__ __forceinline__ void wait_flags(const uint64_t* local_flags, int block, int world, uint64_t epoch) {
  if (threadIdx.x == 0) {
    for (int q = 0; q < world; ++q)
      while (ld_acquire_sys(local_flags + block * world + q) < epoch) __nanosleep(64);
  }
  __syncthreads();
}

// Phase 1: announce "my gradients are complete" to every rank, wait for everybody's announcement, then reduce the
// owned slice over all ranks with peer loads.  grid-stride over the slice; per-CTA partial of the squared norm.
__global__ void __launch_bounds__(PEER_THREADS)
k_peer_reduce(const __grid_constant__ PeerBufs pb, const uint64_t* __restrict__ epoch_ptr, int64_t slice, float grad_scale,
              float* __restrict__ gred, double* __restrict__ cta_partial, unsigned int* __restrict__ ticket) {
  __shared__ double s_red[PEER_THREADS / 32];
  const uint64_t epoch = *epoch_ptr + 1;
  const int W = pb.world, r = pb.rank;
  if (blockIdx.x == 0 && threadIdx.x < W) st_release_sys(pb.flags[threadIdx.x] + 0 * W + r, epoch);
  wait_flags(pb.flags[r], 0, W, epoch);
  const int64_t base = (int64_t)r * slice;
  double acc = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (slice >> 2); i += stride) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < W; ++p) {  // fixed rank order: deterministic
      const float4 g = *reinterpret_cast<const float4*>(pb.grad[p] + base + 4 * i);
      s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
    }
    s.x *= grad_scale; s.y *= grad_scale; s.z *= grad_scale; s.w *= grad_scale;
    reinterpret_cast<float4*>(gred)[i] = s;
    acc += (double)s.x * s.x + (double)s.y * s.y + (double)s.z * s.z + (double)s.w * s.w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < PEER_THREADS / 32; ++w) t += s_red[w];
    cta_partial[blockIdx.x] = t;
    __threadfence();
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) {  // last CTA: this rank's partial norm, published to every rank
      *ticket = 0u;
      __threadfence();
      double n2 = 0.0;
      for (unsigned int c = 0; c < gridDim.x; ++c) n2 += __ldcg(cta_partial + c);
      for (int q = 0; q < W; ++q) {
        pb.norms[q][r] = n2;
        __threadfence_system();
        st_release_sys(pb.flags[q] + 1 * W + r, epoch);
      }
    }
  }
}

// Phase 2: global norm -> clip -> Adam on the owned slice -> parameter slice stored into every rank's buffer.
__global__ void __launch_bounds__(PEER_THREADS)
k_peer_adam(const __grid_constant__ PeerBufs pb, uint64_t* __restrict__ epoch_ptr, int64_t slice,
            const float* __restrict__ gred, float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq, float max_norm,
            float lr, float b1, float b2, float eps, int64_t* __restrict__ step_count, float* __restrict__ norm_out,
            unsigned int* __restrict__ ticket) {
  const uint64_t epoch = *epoch_ptr + 1;
  const int W = pb.world, r = pb.rank;
  wait_flags(pb.flags[r], 1, W, epoch);
  double n2 = 0.0;
  for (int q = 0; q < W; ++q) n2 += pb.norms[r][q];  // same order on every rank -> identical coefficient everywhere
  const float norm = (float)sqrt(n2);
  const float coef = fminf(max_norm / (norm + 1e-6f), 1.0f);
  if (norm_out && blockIdx.x == 0 && threadIdx.x == 0) *norm_out = norm;
  const int64_t step = *step_count + 1;
  const float step_size = (float)((double)lr / (1.0 - pow((double)b1, (double)step)));
  const float inv_bc2_sqrt = (float)(1.0 / sqrt(1.0 - pow((double)b2, (double)step)));
  const int64_t base = (int64_t)r * slice;
  const float* my_param = pb.param[r] + base;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (slice >> 2); i += stride) {
    float4 p = reinterpret_cast<const float4*>(my_param)[i];
    const float4 g = reinterpret_cast<const float4*>(gred)[i];
    float4 m = reinterpret_cast<float4*>(exp_avg)[i], v = reinterpret_cast<float4*>(exp_avg_sq)[i];
    float* pp = &p.x; const float* gg = &g.x; float* mm = &m.x; float* vv = &v.x;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float gc = gg[c] * coef;
      mm[c] = fmaf(gc - mm[c], 1.0f - b1, mm[c]);
      vv[c] = fmaf(vv[c], b2, (1.0f - b2) * gc * gc);
      pp[c] = pp[c] - step_size * (mm[c] / (sqrtf(vv[c]) * inv_bc2_sqrt + eps));
    }
    reinterpret_cast<float4*>(exp_avg)[i] = m;
    reinterpret_cast<float4*>(exp_avg_sq)[i] = v;
    for (int q = 0; q < W; ++q) reinterpret_cast<float4*>(pb.param[q] + base)[i] = p;  // all-gather by peer stores
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) {  // last CTA: slice r is in place on every rank
      *ticket = 0u;
      __threadfence_system();
      for (int q = 0; q < W; ++q) st_release_sys(pb.flags[q] + 2 * W + r, epoch);
      *step_count = step;
    }
  }
}

// Phase 3: every rank's slice has arrived here (so every rank has also finished reading our gradients); advance the epoch.
__global__ void k_peer_fence(const __grid_constant__ PeerBufs pb, uint64_t* __restrict__ epoch_ptr) {
  const uint64_t epoch = *epoch_ptr + 1;
  wait_flags(pb.flags[pb.rank], 2, pb.world, epoch);
  if (threadIdx.x == 0) *epoch_ptr = epoch;
}

}  // namespace

extern "C" {

int rb_peer_scratch_bytes(void) { return (int)(PEER_MAX_CTAS * sizeof(double) + 64); }

int rb_peer_clip_adam(const float* const* peer_grad, float* const* peer_param, uint64_t* const* peer_flags,
                      double* const* peer_norms, int world, int rank, int64_t P, float* gred, float* exp_avg,
                      float* exp_avg_sq, float grad_scale, float max_norm, float lr, float beta1, float beta2, float eps,
                      int64_t* step_count, uint64_t* epoch, void* scratch, float* norm_out, rb_stream_t stream) {
  if (!peer_grad || !peer_param || !peer_flags || !peer_norms || !gred || !exp_avg || !exp_avg_sq || !step_count || !epoch ||
      !scratch)
    return rbi::fail(RB_ERR_INVAL, "rb_peer_clip_adam: null pointer");
  if (world < 1 || world > RB_MAX_PEERS || rank < 0 || rank >= world) return rbi::fail(RB_ERR_RANGE, "rb_peer_clip_adam: bad world/rank");
  if (P <= 0 || P % (4 * (int64_t)world)) return rbi::fail(RB_ERR_INVAL, "rb_peer_clip_adam: P must be a multiple of 4 * world");
  PeerBufs pb;
  pb.world = world;
  pb.rank = rank;
  for (int q = 0; q < RB_MAX_PEERS; ++q) {
    const bool in = q < world;
    if (in && (!peer_grad[q] || !peer_param[q] || !peer_flags[q] || !peer_norms[q])) return rbi::fail(RB_ERR_INVAL, "rb_peer_clip_adam: null peer buffer");
    pb.grad[q] = in ? peer_grad[q] : nullptr;
    pb.param[q] = in ? peer_param[q] : nullptr;
    pb.flags[q] = in ? peer_flags[q] : nullptr;
    pb.norms[q] = in ? peer_norms[q] : nullptr;
  }
  const int64_t slice = P / world;
  int64_t want = (slice / 4 + PEER_THREADS - 1) / PEER_THREADS;
  const int ctas = (int)(want < 1 ? 1 : (want > PEER_MAX_CTAS ? PEER_MAX_CTAS : want));
  double* cta_partial = reinterpret_cast<double*>(scratch);
  unsigned int* tickets = reinterpret_cast<unsigned int*>(cta_partial + PEER_MAX_CTAS);
  cudaStream_t st = (cudaStream_t)stream;
  k_peer_reduce<<<ctas, PEER_THREADS, 0, st>>>(pb, epoch, slice, grad_scale, gred, cta_partial, tickets);
  int rc = rbi::check_launch("rb_peer_clip_adam(reduce)");
  if (rc != RB_OK) return rc;
  k_peer_adam<<<ctas, PEER_THREADS, 0, st>>>(pb, epoch, slice, gred, exp_avg, exp_avg_sq, max_norm, lr, beta1, beta2, eps,
                                            step_count, norm_out, tickets + 1);
  rc = rbi::check_launch("rb_peer_clip_adam(adam)");
  if (rc != RB_OK) return rc;
  k_peer_fence<<<1, 32, 0, st>>>(pb, epoch);
  return rbi::check_launch("rb_peer_clip_adam(fence)");
}

}  // extern "C"
