// rb_kernels.cu -- hand-written sm_100a kernels + C ABI for the Rainbow learner hot path.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC
//        (see rainbow_b200/_build.py).  Header: include/rainbow_b200.h.
//
// Every kernel here is HBM/L2-latency bound byte and index work; none of it is
// GEMM shaped, so there is deliberately no tcgen05/TMEM code in this file.  What matters instead:
// few dependent memory round trips per tree walk, 16-byte coalesced accesses for the frame traffic,
// no host synchronisation anywhere (everything is stream ordered and graph capturable).
//
// Numerics contract: tree nodes are float32 sums recomputed from their
// children (never delta-accumulated), descents carry a float64 residual, the C51 projection uses
// separately rounded float32 operations in the reference's order (no FMA contraction: explicit
// __fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn intrinsics).

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "rainbow_b200.h"
#include "rb_internal.cuh"

namespace rbi {
thread_local char g_err[256] = "";
char* err_buffer() { return g_err; }
bool g_prof_on = false;
ProfKernel g_prof[RB_KERNEL_COUNT];
SmemGrant g_smem_grants[256] = {};
SmemGrant* smem_grants() { return g_smem_grants; }
bool& prof_on() { return g_prof_on; }
ProfKernel* prof_table() { return g_prof; }
}  // namespace rbi

namespace {

using rbi::box_muller;
using rbi::check_launch;
using rbi::fail;
using rbi::normal4;
using rbi::philox4x32_10;
using rbi::ProfKernel;
using rbi::ProfScope;
using rbi::scale_noise;
using rbi::u53;

__device__ __forceinline__ int64_t pymod(int64_t a, int64_t m) {
  int64_t r = a % m;
  return r < 0 ? r + m : r;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__host__ __device__ inline int tree_depth(int64_t tree_start) {  // edges root -> leaf
  int L = 0;
  while (((int64_t)1 << L) - 1 < tree_start) ++L;
  return L;
}

// ================================================================================================
// K4  tree_update : leaf scatter (last write wins) + propagate-to-root, one CTA.
// ================================================================================================
// One thread per updated leaf.  After the leaf writes every thread walks towards the root in
// lock step: a level is  tree[p] = tree[2p+1] + tree[2p+2]  for the thread's parent p, then a CTA
// barrier so the next level reads finished children (threads sharing a parent write the same value).
// Loads bypass L1 (__ldcg): other threads of the CTA have just written those addresses.
constexpr int UPD_THREADS = 1024;

__global__ void __launch_bounds__(UPD_THREADS, 1)
k_tree_update(float* tree, int64_t tree_start, int64_t size, const int64_t* __restrict__ tree_idx,
              const float* __restrict__ raw, float omega, int omega_is_applied, int B, float* running_max,
              int32_t* status, const int32_t* __restrict__ gate) {
  if (gate && *gate == 0) return;   // the batch these priorities belong to was rejected by rb_tree_sample: leave the tree alone
  __shared__ int64_t s_idx[UPD_THREADS];
  __shared__ float s_red[32];
  const int tid = threadIdx.x;
  const int64_t len = tree_start + size;
  float block_max = -CUDART_INF_F;

  for (int base = 0; base < B; base += UPD_THREADS) {
    const int nb = min(UPD_THREADS, B - base);
    int64_t node = -1;
    float val = 0.0f;
    if (tid < nb) {
      node = tree_idx[base + tid];
      float r = raw[base + tid];
      if (omega_is_applied) val = r;
      else if (omega == 0.5f) val = __fsqrt_rn(r);
      else if (omega == 1.0f) val = r;
      else val = (float)pow((double)r, (double)omega);
      if (node < tree_start || node >= len) {
        if (status) atomicExch(status, 1);
        node = -1;
      }
    }
    s_idx[tid] = node;
    __syncthreads();
    if (node >= 0) {
      bool wins = true;
      for (int j = tid + 1; j < nb; ++j) wins &= (s_idx[j] != node);
      if (wins) __stcg(&tree[node], val);
      block_max = fmaxf(block_max, val);
    }
    __syncthreads();
    // every valid thread sits on the leaf level: same number of levels for all
    const int L = tree_depth(tree_start);
    for (int lev = 0; lev < L; ++lev) {
      if (node >= 0) {
        node = (node - 1) >> 1;
        float l = __ldcg(&tree[2 * node + 1]);
        float r = __ldcg(&tree[2 * node + 2]);
        __stcg(&tree[node], __fadd_rn(l, r));
      }
      __syncthreads();
    }
  }
  // running max (memory.py:47-48)
  float m = warp_max(block_max);
  if ((tid & 31) == 0) s_red[tid >> 5] = m;
  __syncthreads();
  if (tid < 32) {
    m = warp_max(s_red[tid]);
    if (tid == 0 && m > *running_max) *running_max = m;
  }
}

// Fast path for B <= 32 (the learner's batch): ONE warp, no barriers.  Every lane first fetches all L siblings
// of its leaf-to-root path in one batch of independent loads (they are untouched by this kernel unless another
// lane's path owns them), then the warp walks the levels in registers: lanes that share a parent find each other
// with __match_any_sync, take the sibling value from the partner lane when the sibling is itself on an updated
// path (else from the prefetched value), and the lowest lane of each group writes the node.  L dependent global
// round trips (one per level, with a CTA barrier each) become one.
constexpr int UPD_MAX_LEVELS = 30;

__global__ void __launch_bounds__(32, 1)
k_tree_update_warp(float* tree, int64_t tree_start, int64_t size, const int64_t* __restrict__ tree_idx,
                   const float* __restrict__ raw, float omega, int omega_is_applied, int B, float* running_max,
                   int32_t* status, const int32_t* __restrict__ gate) {
  if (gate && *gate == 0) return;   // rejected batch (see k_tree_update)
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x;
  const int64_t len = tree_start + size;
  const int L = tree_depth(tree_start);
  int64_t node = -1;
  float val = 0.0f;
  if (lane < B) {
    node = tree_idx[lane];
    float r = raw[lane];
    if (omega_is_applied) val = r;
    else if (omega == 0.5f) val = __fsqrt_rn(r);
    else if (omega == 1.0f) val = r;
    else val = (float)pow((double)r, (double)omega);
    if (node < tree_start || node >= len) {
      if (status) atomicExch(status, 1);
      node = -1;
    }
  }
  const bool active = node >= 0;
  float vmax = active ? val : -CUDART_INF_F;
  vmax = warp_max(vmax);
  // duplicates: the highest lane (= last in index order) wins, memory.py:45 fancy assignment
  {
    const int key = active ? (int)node : -(lane + 1);   // node indices fit 31 bits (checked by the launcher)
    const unsigned grp = __match_any_sync(full, key);
    const int winner = 31 - __clz(grp);
    val = __shfl_sync(full, val, winner);
  }
  // prefetch the siblings along the path (independent loads)
  float sib[UPD_MAX_LEVELS];
#pragma unroll
  for (int l = 0; l < UPD_MAX_LEVELS; ++l) {
    sib[l] = 0.0f;
    if (l < L && active) {
      const int64_t nl = ((node + 1) >> l) - 1;           // ancestor l levels up
      const int64_t sn = (nl & 1) ? nl + 1 : nl - 1;      // odd index = left child
      sib[l] = __ldcg(tree + sn);
    }
  }
  {
    const int key = active ? (int)node : -(lane + 1);
    const unsigned grp = __match_any_sync(full, key);
    if (active && lane == __ffs(grp) - 1) __stcg(tree + node, val);
  }
#pragma unroll
  for (int l = 0; l < UPD_MAX_LEVELS; ++l) {
    if (l < L) {
      const int64_t parent = active ? ((node - 1) >> 1) : -(int64_t)(lane + 1);
      const bool is_left = active && (node & 1);
      const unsigned grp = __match_any_sync(full, (int)parent);
      const unsigned lefts = __ballot_sync(full, is_left);
      const unsigned other = grp & (is_left ? ~lefts : lefts);
      const float partner = __shfl_sync(full, val, other ? __ffs(other) - 1 : lane);
      const float sv = other ? partner : sib[l];
      val = __fadd_rn(val, sv);                            // fl32(left + right); float addition commutes
      node = parent;
      if (active && lane == __ffs(grp) - 1) __stcg(tree + node, val);
    }
  }
  if (lane == 0 && vmax > *running_max) *running_max = vmax;   // memory.py:47-48
}

// ================================================================================================
// Tree descent shared by K1 (sample) and rb_tree_find.
// ================================================================================================
// A warp walks one sample.  The top TOP_DEPTH levels are staged in shared memory by the CTA; below
// that the warp fetches up to five levels of the current subtree at once: relative to node i the
// nodes of depth j form the contiguous run [(i+1)*2^j - 1, (i+1)*2^j - 1 + 2^j), so lanes 2..31
// fetch depths 1..4 (30 nodes) and all 32 lanes fetch depth 5 as one 128-byte line.  The five
// compare/subtract steps then run on register values exchanged with shuffles: one memory round
// trip per five levels instead of five.
constexpr int TOP_DEPTH = 10;                        // depths 0..10 staged: 2047 nodes, 8 KB
constexpr int TOP_NODES = (2 << TOP_DEPTH) - 1;

struct Descent {
  int64_t node;
  float prob;
};

__device__ __forceinline__ Descent warp_descend(const float* __restrict__ tree, const float* s_top, int64_t tree_start,
                                                int64_t len, int L, double v) {
  const int lane = threadIdx.x & 31;
  int64_t i = 0;
  int d = 0;
  const int S = min(L, TOP_DEPTH);
  // ---- shared-memory phase (all lanes redundantly, uniform) ----
  for (; d < S; ++d) {
    int64_t cl = 2 * i + 1, cr = cl + 1;
    if (cl >= tree_start) {  // children are leaves: clip like memory.py:70-71
      cl = min(cl, len - 1);
      cr = min(cr, len - 1);
    }
    float left = s_top[cl];
    bool right = v > (double)left;
    if (right) v = __dsub_rn(v, (double)left);
    i = right ? cr : cl;
  }
  float prob = (L <= TOP_DEPTH) ? s_top[i] : 0.0f;
  // ---- chunked phase ----
  while (d < L) {
    const int nl = min(5, L - d);
    float lo = 0.0f, hi = 0.0f;
    {
      int h = lane;  // subtree positions 2..31 -> depths 1..4
      if (h >= 2) {
        int j = 31 - __clz(h);
        if (j <= nl) {
          int64_t nd = (i << j) + h - 1;
          if (d + j == L) nd = min(nd, len - 1);
          lo = __ldg(tree + nd);
        }
      }
      h = lane + 32;  // positions 32..63 -> depth 5
      if (nl == 5) {
        int64_t nd = (i << 5) + h - 1;
        if (d + 5 == L) nd = min(nd, len - 1);
        hi = __ldg(tree + nd);
      }
    }
    int h = 1;
    for (int j = 0; j < nl; ++j) {
      int hl = 2 * h;
      float a = __shfl_sync(0xffffffffu, lo, hl & 31);
      float b = __shfl_sync(0xffffffffu, hi, hl & 31);
      float left = (hl < 32) ? a : b;
      bool right = v > (double)left;
      if (right) v = __dsub_rn(v, (double)left);
      h = hl + (right ? 1 : 0);
    }
    {
      float a = __shfl_sync(0xffffffffu, lo, h & 31);
      float b = __shfl_sync(0xffffffffu, hi, h & 31);
      prob = (h < 32) ? a : b;
    }
    i = (i << nl) + h - 1;
    d += nl;
    if (d == L) i = min(i, len - 1);
  }
  Descent r;
  r.node = i;
  r.prob = prob;
  return r;
}

__device__ __forceinline__ void stage_top(const float* __restrict__ tree, float* s_top, int64_t len) {
  const int ntop = (int)min((int64_t)TOP_NODES, len);
  for (int i = threadIdx.x; i < ntop; i += blockDim.x) s_top[i] = __ldg(tree + i);
  __syncthreads();
}

constexpr int SAMPLE_THREADS = 1024;

__global__ void __launch_bounds__(SAMPLE_THREADS, 1)
k_tree_find(const float* __restrict__ tree, int64_t tree_start, int64_t size, const double* __restrict__ values,
            int B, float* probs, int64_t* data_idx, int64_t* tree_idx) {
  __shared__ float s_top[TOP_NODES];
  const int64_t len = tree_start + size;
  const int L = tree_depth(tree_start);
  stage_top(tree, s_top, len);
  const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, lane = threadIdx.x & 31;
  for (int k = blockIdx.x * nwarps + warp; k < B; k += gridDim.x * nwarps) {
    Descent r = warp_descend(tree, s_top, tree_start, len, L, values[k]);
    if (lane == 0) {
      probs[k] = r.prob;
      data_idx[k] = r.node - tree_start;
      tree_idx[k] = r.node;
    }
  }
}

// ================================================================================================
// K1  tree_sample : stratified proportional sampling with whole-batch rejection + IS weights.
// ================================================================================================
__global__ void __launch_bounds__(SAMPLE_THREADS, 1)
k_tree_sample(const float* __restrict__ tree, int64_t tree_start, int64_t size, const int64_t* __restrict__ ring_state,
              int n, int history, const double* __restrict__ u01, int u01_attempts, uint64_t seed,
              unsigned long long* rng_counter, int B, float beta, const float* __restrict__ beta_dev, int max_attempts,
              float* probs, int64_t* data_idx, int64_t* tree_idx, float* weights, int32_t* status) {
  __shared__ float s_top[TOP_NODES];
  __shared__ int s_valid;
  __shared__ float s_red[32];
  __shared__ float s_prob[SAMPLE_THREADS];   // leaf values of the first SAMPLE_THREADS samples (kept on chip for the weights)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  const int64_t len = tree_start + size;
  const int L = tree_depth(tree_start);
  if (tid == 0) s_valid = 1;
  // every scalar the kernel needs is requested before the staging barrier, so these round trips overlap the staging loads
  const int64_t head = ring_state[0];
  const bool full = ring_state[1] != 0;
  const unsigned long long ctr0 = (u01 == nullptr) ? *rng_counter : 0ull;
  const float b = beta_dev ? *beta_dev : beta;
  stage_top(tree, s_top, len);

  const float p_total = s_top[0];                          // memory.py:149
  const float seg = __fdiv_rn(p_total, (float)B);          // memory.py:125 (float32)
  const double segd = (double)seg;
  const int tries = (u01 != nullptr) ? u01_attempts : max_attempts;

  int attempt = 0;
  bool ok_batch = false;
  for (; attempt < tries; ++attempt) {
    for (int k = warp; k < B; k += nwarps) {
      double u;
      if (u01 != nullptr) {
        u = u01[(size_t)attempt * B + k];
      } else {
        unsigned long long c = ctr0 + (unsigned long long)attempt;
        uint4 r = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)k, 0x5A4D504Cu),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        u = u53(r.x, r.y);
      }
      // memory.py:126,129: uniform(0, seg) + k*seg, float64, separately rounded
      double v = __dadd_rn(__dadd_rn(0.0, __dmul_rn(segd, u)), __dmul_rn((double)k, segd));
      Descent r = warp_descend(tree, s_top, tree_start, len, L, v);
      if (lane == 0) {
        int64_t di = r.node - tree_start;
        probs[k] = r.prob;
        if (k < SAMPLE_THREADS) s_prob[k] = r.prob;
        data_idx[k] = di;
        tree_idx[k] = r.node;
        bool ok = (pymod(head - di, size) > (int64_t)n) && (pymod(di - head, size) >= (int64_t)history) &&
                  (r.prob != 0.0f);                         // memory.py:131
        if (!ok) s_valid = 0;
      }
    }
    __syncthreads();
    ok_batch = (s_valid != 0);
    __syncthreads();
    if (ok_batch) { ++attempt; break; }
    if (tid == 0) s_valid = 1;
    __syncthreads();
  }

  // importance-sampling weights (memory.py:151-154), float32 like numpy
  const float nb = -b;
  const float count = (float)(full ? size : head);
  float wmax = -CUDART_INF_F;
  float w_first = 0.0f;   // weight of sample k = tid (kept in a register), later samples go through global memory
  for (int k = tid; k < B; k += blockDim.x) {
    float p = __fdiv_rn(k < SAMPLE_THREADS ? s_prob[k] : probs[k], p_total);
    float w = (float)pow((double)__fmul_rn(count, p), (double)nb);
    if (k == tid) w_first = w; else weights[k] = w;
    wmax = fmaxf(wmax, w);
  }
  wmax = warp_max(wmax);
  if (lane == 0) s_red[warp] = wmax;
  __syncthreads();
  wmax = warp_max(s_red[lane]);
  // A batch that is still invalid after the last allowed redraw (the reference would keep redrawing, memory.py:128-132)
  // must not train anything: its weights are zeroed -- so the loss gradient is exactly zero -- and status[0] = 0 gates
  // the optimiser step and the priority write-back (rb_clip_adam / rb_tree_update `gate`).
  for (int k = tid; k < B; k += blockDim.x)
    weights[k] = ok_batch ? __fdiv_rn(k == tid ? w_first : weights[k], wmax) : 0.0f;
  if (tid == 0) {
    status[0] = ok_batch ? 1 : 0;
    status[1] = attempt;
    if (!ok_batch && u01 == nullptr) status[2] = status[2] + 1;   // batches rejected for good so far (host diagnostics)
    if (u01 == nullptr) *rng_counter = ctr0 + (unsigned long long)attempt;
  }
}

// ================================================================================================
// K2  gather : frame-stack + n-step window gather with episode-boundary blanking.
// ================================================================================================
// grid = (used_slots * split, B).  A CTA converts (a slice of) ONE stored uint8 frame and writes it
// to every place it appears in the outputs (state slot and/or next-state slot), so each stored frame
// is read once.  16-byte loads, 4 x 16-byte stores per thread-iteration.
constexpr int GATHER_THREADS = 256;
constexpr int FRAME_VEC = RB_FRAME_BYTES / 16;  // 441 uint4 per frame

__device__ __forceinline__ uint64_t window_first_bits(const int32_t* __restrict__ timestep, int64_t size, int64_t idx,
                                                      int history, int W) {
  // lane s (and s+32) looks at window record s; ballot -> 64-bit "timestep == 0" mask (memory.py:114)
  const int lane = threadIdx.x & 31;
  bool f0 = false, f1 = false;
  if (lane < W) f0 = __ldg(timestep + pymod(idx - (history - 1) + lane, size)) == 0;
  if (lane + 32 < W) f1 = __ldg(timestep + pymod(idx - (history - 1) + lane + 32, size)) == 0;
  uint32_t b0 = __ballot_sync(0xffffffffu, f0), b1 = __ballot_sync(0xffffffffu, f1);
  return (uint64_t)b0 | ((uint64_t)b1 << 32);
}

__host__ __device__ __forceinline__ uint64_t low_bits(int k) {  // k in [0,64]
  return k >= 64 ? ~0ull : ((1ull << k) - 1ull);
}

// memory.py:116-119: slot s < H-1 is blank if any of s+1..H-1 starts an episode; slot s >= H is blank
// if any of H..s starts one; slot H-1 never is.
__device__ __forceinline__ bool slot_blank(uint64_t first, int s, int history) {
  if (s < history - 1) return (first & low_bits(history) & ~low_bits(s + 1)) != 0;
  if (s >= history) return (first & low_bits(s + 1) & ~low_bits(history)) != 0;
  return false;
}

__device__ __forceinline__ float4 u8x4_to_unit(uint32_t w) {
  return make_float4(__fdiv_rn((float)(w & 0xffu), 255.0f), __fdiv_rn((float)((w >> 8) & 0xffu), 255.0f),
                     __fdiv_rn((float)((w >> 16) & 0xffu), 255.0f), __fdiv_rn((float)(w >> 24), 255.0f));
}

__global__ void __launch_bounds__(GATHER_THREADS)
k_gather(const uint8_t* __restrict__ frames, const int32_t* __restrict__ timestep, const int32_t* __restrict__ action,
         const float* __restrict__ reward, const uint8_t* __restrict__ nonterminal, int64_t size,
         const int64_t* __restrict__ data_idx, int B, int history, int n, const float* __restrict__ gamma_pow,
         float* __restrict__ states, float* __restrict__ next_states, int64_t* __restrict__ actions,
         float* __restrict__ returns, float* __restrict__ nonterminals, int split) {
  __shared__ uint64_t s_first;
  const int b = blockIdx.y;
  const int W = history + n;
  const int used = blockIdx.x / split, part = blockIdx.x % split;
  // used-slot -> window slot: slots [0,H) feed `states`, [n,n+H) feed `next_states`
  const int s = (n >= history && used >= history) ? n + (used - history) : used;
  const int64_t idx = data_idx[b];
  const int64_t pos = pymod(idx - (history - 1) + s, size);
  const int per = (FRAME_VEC + split - 1) / split;
  const int v0 = part * per, v1 = min(FRAME_VEC, v0 + per);
  // issue the (rarely discarded) frame load first so it overlaps the timestep loads that decide the blanking
  const uint4* src = reinterpret_cast<const uint4*>(frames + (size_t)pos * RB_FRAME_BYTES);
  uint4 pre = make_uint4(0, 0, 0, 0);
  if (v0 + (int)threadIdx.x < v1) pre = __ldg(src + v0 + threadIdx.x);
  if (threadIdx.x < 32) {
    uint64_t f = window_first_bits(timestep, size, idx, history, W);
    if (threadIdx.x == 0) s_first = f;
  }
  __syncthreads();
  const uint64_t first = s_first;
  const bool blank = slot_blank(first, s, history);

  float4* dst_s = (s < history) ? reinterpret_cast<float4*>(states + ((size_t)b * history + s) * RB_FRAME_BYTES) : nullptr;
  float4* dst_n = (s >= n && s < n + history)
                      ? reinterpret_cast<float4*>(next_states + ((size_t)b * history + (s - n)) * RB_FRAME_BYTES)
                      : nullptr;
  for (int v = v0 + threadIdx.x; v < v1; v += GATHER_THREADS) {
    uint4 q = blank ? make_uint4(0, 0, 0, 0) : (v == v0 + (int)threadIdx.x ? pre : __ldg(src + v));
    float4 a = u8x4_to_unit(q.x), bq = u8x4_to_unit(q.y), c = u8x4_to_unit(q.z), d = u8x4_to_unit(q.w);
    if (dst_s) {
      __stcs(dst_s + 4 * v + 0, a); __stcs(dst_s + 4 * v + 1, bq); __stcs(dst_s + 4 * v + 2, c); __stcs(dst_s + 4 * v + 3, d);
    }
    if (dst_n) {
      __stcs(dst_n + 4 * v + 0, a); __stcs(dst_n + 4 * v + 1, bq); __stcs(dst_n + 4 * v + 2, c); __stcs(dst_n + 4 * v + 3, d);
    }
  }
  // per-sample scalars, once per sample (memory.py:140-145)
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const int sa = history - 1;
    actions[b] = (int64_t)__ldg(action + pymod(idx, size));  // slot H-1 is never blanked
    float acc = 0.0f;
    for (int k = 0; k < n; ++k) {
      int sk = sa + k;
      float r = slot_blank(first, sk, history) ? 0.0f : __ldg(reward + pymod(idx + k, size));
      acc = __fadd_rn(acc, __fmul_rn(r, __ldg(gamma_pow + k)));
    }
    returns[b] = acc;
    const int sl = W - 1;
    nonterminals[b] = slot_blank(first, sl, history) ? 0.0f : (__ldg(nonterminal + pymod(idx + n, size)) ? 1.0f : 0.0f);
  }
}

// memory.py:166-178 validation iterator, batched: grid = (history, count)
__global__ void __launch_bounds__(GATHER_THREADS)
k_iter_states(const uint8_t* __restrict__ frames, const int32_t* __restrict__ timestep, int64_t size, int64_t first_idx,
              int history, float* __restrict__ out) {
  __shared__ uint64_t s_first;
  const int s = blockIdx.x;
  const int64_t cur = first_idx + blockIdx.y;
  if (threadIdx.x < 32) {
    uint64_t f = window_first_bits(timestep, size, cur, history, history);
    if (threadIdx.x == 0) s_first = f;
  }
  __syncthreads();
  const bool blank = slot_blank(s_first, s, history);
  const int64_t pos = pymod(cur - (history - 1) + s, size);
  const uint4* src = reinterpret_cast<const uint4*>(frames + (size_t)pos * RB_FRAME_BYTES);
  float4* dst = reinterpret_cast<float4*>(out + ((size_t)blockIdx.y * history + s) * RB_FRAME_BYTES);
  for (int v = threadIdx.x; v < FRAME_VEC; v += GATHER_THREADS) {
    uint4 q = blank ? make_uint4(0, 0, 0, 0) : __ldg(src + v);
    dst[4 * v + 0] = u8x4_to_unit(q.x);
    dst[4 * v + 1] = u8x4_to_unit(q.y);
    dst[4 * v + 2] = u8x4_to_unit(q.z);
    dst[4 * v + 3] = u8x4_to_unit(q.w);
  }
}

// ================================================================================================
// K5  append : quantise + store one transition, set its leaf to the running max, walk to the root.
// ================================================================================================
constexpr int APPEND_THREADS = 256;

__global__ void __launch_bounds__(APPEND_THREADS)
k_append(float* tree, int64_t tree_start, int64_t size, uint8_t* __restrict__ frames, int32_t* timestep,
         int32_t* action, float* reward, uint8_t* nonterminal, int64_t* ring_state, const float* running_max,
         const float* __restrict__ state_last, int32_t action_value, float reward_value, int terminal) {
  const int64_t head = ring_state[0];
  const int64_t t_ep = ring_state[2];
  const int tid = threadIdx.x;
  // frame: f32 * 255 then truncating cast (memory.py:106); 4 pixels per thread-iteration
  uint32_t* dst = reinterpret_cast<uint32_t*>(frames + (size_t)head * RB_FRAME_BYTES);
  const float4* src = reinterpret_cast<const float4*>(state_last);
  {  // all of a thread's loads in flight at once: the frame may sit in pinned host memory (a PCIe round trip per load)
    constexpr int PER = (RB_FRAME_BYTES / 4 + APPEND_THREADS - 1) / APPEND_THREADS;   // 7
    float4 x[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int v = tid + u * APPEND_THREADS;
      x[u] = (v < RB_FRAME_BYTES / 4) ? src[v] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int v = tid + u * APPEND_THREADS;
      if (v < RB_FRAME_BYTES / 4) {
        uint32_t a = (uint32_t)(uint8_t)(int)__fmul_rn(x[u].x, 255.0f);
        uint32_t b = (uint32_t)(uint8_t)(int)__fmul_rn(x[u].y, 255.0f);
        uint32_t c = (uint32_t)(uint8_t)(int)__fmul_rn(x[u].z, 255.0f);
        uint32_t d = (uint32_t)(uint8_t)(int)__fmul_rn(x[u].w, 255.0f);
        dst[v] = a | (b << 8) | (c << 16) | (d << 24);
      }
    }
  }
  if (tid < 32) {
    // warp 0: the walk.  All siblings along the path are independent of the new value, so lane j
    // fetches the sibling at level j (one round trip), then lane 0 chains the float32 additions
    // child + sibling (commutative, so left/right order does not matter bit-wise).
    const int L = tree_depth(tree_start);
    const float value = *running_max;  // memory.py:107: new transitions get the max priority
    const int64_t leaf = head + tree_start;
    float sib = 0.0f;
    if (tid < L) {
      int64_t node = leaf;
      for (int j = 0; j < tid; ++j) node = (node - 1) >> 1;
      int64_t sibling = (node & 1) ? node + 1 : node - 1;  // odd = left child
      sib = __ldcg(tree + sibling);
    }
    float acc = value;
    int64_t node = leaf;
    if (tid == 0) __stcg(tree + node, acc);
    for (int j = 0; j < L; ++j) {
      float sj = __shfl_sync(0xffffffffu, sib, j);
      acc = __fadd_rn(acc, sj);
      node = (node - 1) >> 1;
      if (tid == 0) __stcg(tree + node, acc);
    }
    if (tid == 0) {
      timestep[head] = (int32_t)t_ep;
      action[head] = action_value;
      reward[head] = reward_value;
      nonterminal[head] = terminal ? 0 : 1;
    }
  }
  __syncthreads();
  if (tid == 0) {
    int64_t nh = head + 1 == size ? 0 : head + 1;
    ring_state[0] = nh;
    if (nh == 0) ring_state[1] = 1;
    ring_state[2] = terminal ? 0 : t_ep + 1;
    ring_state[3] = ring_state[3] + 1;
  }
}

// ------------------------------------------------------------------------------------------------
// K5b  append_batch : up to RB_APPEND_BATCH queued transitions in ONE launch (actor side).
// ------------------------------------------------------------------------------------------------
// Equivalent to calling k_append k times in order: the records go to slots head, head+1, ... (mod size), the
// in-episode counter follows the terminals, every new leaf gets the running max (which appends never change) and,
// because every internal node is recomputed from its children, the tree after k sequential walks equals the tree
// after ONE level-synchronous batched walk over the k leaves (same argument as for rb_tree_update).  The frame
// pointers may be device memory or pinned host memory (read in place over PCIe: no staging copy, no extra launch).
struct AppendBatch {
  const float* frame[RB_APPEND_BATCH];
  int32_t action[RB_APPEND_BATCH];
  float reward[RB_APPEND_BATCH];
  int32_t terminal[RB_APPEND_BATCH];
  int k;
};

__global__ void __launch_bounds__(APPEND_THREADS)
k_append_batch(float* tree, int64_t tree_start, int64_t size, uint8_t* __restrict__ frames, int32_t* timestep,
               int32_t* action, float* reward, uint8_t* nonterminal, int64_t* ring_state, const float* running_max,
               const __grid_constant__ AppendBatch ab) {
  // grid = k CTAs: CTA j quantises frame j (all of a thread's 16-byte loads in flight at once -- the frames may sit in
  // pinned host memory, where every dependent load is a PCIe round trip); warp 0 of CTA 0 also writes the k records and
  // walks the tree.  ring_state is read by every CTA when it starts and advanced by the LAST CTA to finish (ticket in
  // ring_state[4]), so no CTA can see the new head.
  const int tid = threadIdx.x;
  const int j = blockIdx.x;
  const int64_t head = ring_state[0];
  const int64_t t_ep0 = ring_state[2];
  const int k = ab.k;
  {  // frame j: f32 * 255 then truncating cast (memory.py:106)
    int64_t slot = head + j;
    if (slot >= size) slot -= size;
    uint32_t* dst = reinterpret_cast<uint32_t*>(frames + (size_t)slot * RB_FRAME_BYTES);
    const float4* src = reinterpret_cast<const float4*>(ab.frame[j]);
    constexpr int PER = (RB_FRAME_BYTES / 4 + APPEND_THREADS - 1) / APPEND_THREADS;   // 7
    float4 x[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int v = tid + u * APPEND_THREADS;
      x[u] = (v < RB_FRAME_BYTES / 4) ? src[v] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int v = tid + u * APPEND_THREADS;
      if (v < RB_FRAME_BYTES / 4) {
        const uint32_t a = (uint32_t)(uint8_t)(int)__fmul_rn(x[u].x, 255.0f);
        const uint32_t b = (uint32_t)(uint8_t)(int)__fmul_rn(x[u].y, 255.0f);
        const uint32_t c = (uint32_t)(uint8_t)(int)__fmul_rn(x[u].z, 255.0f);
        const uint32_t d = (uint32_t)(uint8_t)(int)__fmul_rn(x[u].w, 255.0f);
        dst[v] = a | (b << 8) | (c << 16) | (d << 24);
      }
    }
  }
  if (j == 0 && tid < 32) {
    const unsigned full = 0xffffffffu;
    const int lane = tid;
    const int L = tree_depth(tree_start);
    const bool active = lane < k;
    int64_t slot = head + lane;
    if (slot >= size) slot -= size;
    int64_t node = active ? slot + tree_start : -1;
    float val = *running_max;  // memory.py:107
    if (active) {              // record fields; timestep follows the terminals of the earlier queued transitions
      int64_t t = t_ep0;
      for (int i = 0; i < lane; ++i) t = ab.terminal[i] ? 0 : t + 1;
      timestep[slot] = (int32_t)t;
      action[slot] = ab.action[lane];
      reward[slot] = ab.reward[lane];
      nonterminal[slot] = ab.terminal[lane] ? 0 : 1;
    }
    float sib[32];
#pragma unroll
    for (int l = 0; l < 32; ++l) {
      sib[l] = 0.0f;
      if (l < L && active) {
        const int64_t nl = ((node + 1) >> l) - 1;
        const int64_t sn = (nl & 1) ? nl + 1 : nl - 1;
        sib[l] = __ldcg(tree + sn);
      }
    }
    if (active) __stcg(tree + node, val);  // k distinct leaves (k <= size is checked by the launcher)
#pragma unroll
    for (int l = 0; l < 32; ++l) {
      if (l < L) {
        const int64_t parent = active ? ((node - 1) >> 1) : -(int64_t)(lane + 1);
        const bool is_left = active && (node & 1);
        const unsigned grp = __match_any_sync(full, (int)parent);
        const unsigned lefts = __ballot_sync(full, is_left);
        const unsigned other = grp & (is_left ? ~lefts : lefts);
        const float partner = __shfl_sync(full, val, other ? __ffs(other) - 1 : lane);
        val = __fadd_rn(val, other ? partner : sib[l]);
        node = parent;
        if (active && lane == __ffs(grp) - 1) __stcg(tree + node, val);
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    unsigned long long* ticket = reinterpret_cast<unsigned long long*>(ring_state + 4);
    if (atomicAdd(ticket, 1ull) == (unsigned long long)(k - 1)) {   // last CTA: everybody has read the old head
      *ticket = 0ull;
      int64_t nh = head + k;
      bool wrapped = false;
      if (nh >= size) {
        nh -= size;
        wrapped = true;
      }
      int64_t t = t_ep0;
      for (int i = 0; i < k; ++i) t = ab.terminal[i] ? 0 : t + 1;
      ring_state[0] = nh;
      if (wrapped) ring_state[1] = 1;
      ring_state[2] = t;
      ring_state[3] = ring_state[3] + k;
    }
  }
}

// ================================================================================================
// K3  c51_loss_grad : double-DQN argmax + categorical projection + IS-weighted CE loss + gradient.
// ================================================================================================
// One warp per sample; lane owns atoms z = lane + 32*r.  The projected distribution is built as a
// GATHER (thread per target atom scans the source atoms in order, l-side terms first, then u-side
// terms): deterministic and in the reference CPU index_add_ order (agent.py:91-92), no atomics.
// c51_core works on logit ROWS given by generic pointers (global memory for the plain entry point,
// warp-private shared memory for the dueling entry point) and returns, per lane, the gradient row
// g[z] = (w/B)(p*sum(m) - m) of the taken action.
constexpr int C51_WARPS = 4;
// atoms per lane: templates are instantiated for R = 2 (Z <= 64, the usual 51 atoms) and R = 4 (Z <= 128)

struct C51Scratch {  // per warp
  float pt[RB_MAX_ATOMS];  // target probabilities p(s', a*)
  float b[RB_MAX_ATOMS];
  int l[RB_MAX_ATOMS];
  int u[RB_MAX_ATOMS];
};

template <int C51_R>
__device__ __forceinline__ void softmax_row(const float* row, int Z, int lane, float (&e)[C51_R], float (&x)[C51_R],
                                            float& mx, float& sum) {
  mx = -CUDART_INF_F;
#pragma unroll
  for (int r = 0; r < C51_R; ++r) {
    int z = lane + 32 * r;
    x[r] = (z < Z) ? row[z] : -CUDART_INF_F;
    mx = fmaxf(mx, x[r]);
  }
  mx = warp_max(mx);
  sum = 0.0f;
#pragma unroll
  for (int r = 0; r < C51_R; ++r) {
    e[r] = (lane + 32 * r < Z) ? expf(x[r] - mx) : 0.0f;
    sum = __fadd_rn(sum, e[r]);
  }
  sum = warp_sum(sum);
}

// Expected value sum_z support_z * p_z of one logit row held in registers (x[r] = logit of atom lane + 32 r, -inf past Z),
// p = e / sum(e): the numerator and sum(e) ride the same shuffle butterfly (two independent shuffles per step) and one
// division finishes the row (agent.py:71-72).  All lanes return the value.
template <int C51_R>
__device__ __forceinline__ float c51_expected_value(const float (&x)[C51_R], const float (&sup)[C51_R], int Z, int lane) {
  float mx = -CUDART_INF_F;
#pragma unroll
  for (int r = 0; r < C51_R; ++r) mx = fmaxf(mx, x[r]);
  mx = warp_max(mx);
  float se = 0.0f, sn = 0.0f;
#pragma unroll
  for (int r = 0; r < C51_R; ++r) {
    const float ee = (lane + 32 * r < Z) ? expf(x[r] - mx) : 0.0f;
    se = __fadd_rn(se, ee);
    sn = __fadd_rn(sn, __fmul_rn(sup[r], ee));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    se = __fadd_rn(se, __shfl_xor_sync(0xffffffffu, se, o));
    sn = __fadd_rn(sn, __shfl_xor_sync(0xffffffffu, sn, o));
  }
  return __fdiv_rn(sn, se);
}

// q_on_ns / q_tg_ns: A rows of Z (row stride Z); q_on_s_act: the row of the taken action.
// best_known >= 0: a* was already determined by the caller (q_on_ns is then not read, q_tg_ns points at the row of a*).
template <int C51_R>
__device__ __forceinline__ void c51_core(C51Scratch& sc, int lane, int i, int B, int A, int Z, const float* q_on_ns,
                                         const float* q_tg_ns, const float* q_on_s_act, float ret, float nonterminal,
                                         float weight, const float* __restrict__ support, float vmin, float vmax,
                                         float delta_z, float gamma_n, float* __restrict__ loss, float* __restrict__ m_out,
                                         int64_t* __restrict__ astar_out, float (&g)[C51_R], int best_known = -1) {
  float sup[C51_R];
#pragma unroll
  for (int r = 0; r < C51_R; ++r) {
    int z = lane + 32 * r;
    sup[r] = (z < Z) ? __ldg(support + z) : 0.0f;
  }
  float e[C51_R], x[C51_R], mx, sum;

  // ---- agent.py:71-73: a* = argmax_a sum_z support_z * softmax(q_online(s'))[a,z] ----
  int best = 0;
  if (best_known >= 0) {
    best = best_known;
  } else {
    float best_ev = -CUDART_INF_F;
    for (int a = 0; a < A; ++a) {
      const float* row = q_on_ns + (size_t)a * Z;
#pragma unroll
      for (int r = 0; r < C51_R; ++r) {
        int z = lane + 32 * r;
        x[r] = (z < Z) ? row[z] : -CUDART_INF_F;
      }
      const float ev = c51_expected_value<C51_R>(x, sup, Z, lane);
      if (ev > best_ev) {  // first maximum wins, like torch.argmax
        best_ev = ev;
        best = a;
      }
    }
  }
  if (astar_out && lane == 0) astar_out[i] = best;

  // ---- agent.py:75-76: target distribution of the selected action ----
  softmax_row(q_tg_ns + (best_known >= 0 ? (size_t)0 : (size_t)best * Z), Z, lane, e, x, mx, sum);
#pragma unroll
  for (int r = 0; r < C51_R; ++r) {
    int z = lane + 32 * r;
    if (z < Z) sc.pt[z] = __fdiv_rn(e[r], sum);
  }

  // ---- agent.py:66-67: log p(s, a) and p(s, a) of the online net ----
  float p_on[C51_R], logp[C51_R];
  softmax_row(q_on_s_act, Z, lane, e, x, mx, sum);
  {
    const float lsum = logf(sum);
#pragma unroll
    for (int r = 0; r < C51_R; ++r) {
      p_on[r] = __fdiv_rn(e[r], sum);
      logp[r] = (lane + 32 * r < Z) ? __fsub_rn(__fsub_rn(x[r], mx), lsum) : 0.0f;
    }
  }

  // ---- agent.py:79-86: Tz, b, l, u ----
  const float scale = __fmul_rn(nonterminal, gamma_n);
#pragma unroll
  for (int r = 0; r < C51_R; ++r) {
    int z = lane + 32 * r;
    if (z < Z) {
      float tz = __fadd_rn(ret, __fmul_rn(scale, sup[r]));
      tz = fminf(fmaxf(tz, vmin), vmax);
      float bb = __fdiv_rn(__fsub_rn(tz, vmin), delta_z);
      int lo = (int)floorf(bb), up = (int)ceilf(bb);
      if (up > 0 && lo == up) lo -= 1;
      if (lo < Z - 1 && lo == up) up += 1;
      sc.b[z] = bb;
      sc.l[z] = lo;
      sc.u[z] = up;
    }
  }
  __syncwarp();

  // ---- agent.py:89-92: m, deterministic gather in index_add_ order ----
  // Tz is non-decreasing in the atom index (support increasing, scale >= 0) and u == l + 1 after the fix-ups, so
  // the source atoms feeding target k on the l side form the contiguous run {j : l[j] == k} and on the u side the
  // run {j : l[j] == k - 1}: two binary searches replace the 2*Z-long scans.  The additions still happen in atom
  // order, l side first (bit-identical to the scan).  Odd inputs (NaN, decreasing support) take the full scan.
  bool mono = true;
#pragma unroll
  for (int r = 0; r < C51_R; ++r) {
    const int z = lane + 32 * r;
    if (z < Z) mono = mono && (sc.u[z] == sc.l[z] + 1) && (z + 1 >= Z || sc.l[z] <= sc.l[z + 1]);
  }
  mono = __all_sync(0xffffffffu, mono);
  float m[C51_R];
  float ce = 0.0f, msum = 0.0f;
#pragma unroll
  for (int r = 0; r < C51_R; ++r) {
    const int k = lane + 32 * r;
    float acc = 0.0f;
    if (k < Z) {
      if (mono) {
        int lo = 0, hi = Z;              // first j with l[j] >= k - 1
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (sc.l[mid] < k - 1) lo = mid + 1; else hi = mid;
        }
        int j = lo;
        const int ju = j;                // start of the u-side run (l == k - 1)
        while (j < Z && sc.l[j] == k - 1) ++j;
        for (int jj = j; jj < Z && sc.l[jj] == k; ++jj)
          acc = __fadd_rn(acc, __fmul_rn(sc.pt[jj], __fsub_rn((float)sc.u[jj], sc.b[jj])));
        for (int jj = ju; jj < j; ++jj)
          acc = __fadd_rn(acc, __fmul_rn(sc.pt[jj], __fsub_rn(sc.b[jj], (float)sc.l[jj])));
      } else {
        for (int j = 0; j < Z; ++j)
          if (sc.l[j] == k) acc = __fadd_rn(acc, __fmul_rn(sc.pt[j], __fsub_rn((float)sc.u[j], sc.b[j])));
        for (int j = 0; j < Z; ++j)
          if (sc.u[j] == k) acc = __fadd_rn(acc, __fmul_rn(sc.pt[j], __fsub_rn(sc.b[j], (float)sc.l[j])));
      }
      if (m_out) m_out[(size_t)i * Z + k] = acc;
    }
    m[r] = acc;
    ce = __fadd_rn(ce, __fmul_rn(acc, logp[r]));
    msum = __fadd_rn(msum, acc);
  }
  ce = warp_sum(ce);
  msum = warp_sum(msum);
  if (lane == 0) loss[i] = -ce;  // agent.py:94

  // ---- agent.py:96: d mean(w*loss) / d q_online(s)[i, act, :] ----
  const float wi = __fdiv_rn(weight, (float)B);
#pragma unroll
  for (int r = 0; r < C51_R; ++r) g[r] = __fmul_rn(wi, __fsub_rn(__fmul_rn(p_on[r], msum), m[r]));
  __syncwarp();
}

template <int C51_R>
__global__ void __launch_bounds__(C51_WARPS * 32)
k_c51(const float* __restrict__ q_on_s, const float* __restrict__ q_on_ns, const float* __restrict__ q_tg_ns,
      const int64_t* __restrict__ actions, const float* __restrict__ returns, const float* __restrict__ nonterminals,
      const float* __restrict__ weights, const float* __restrict__ support, float vmin, float vmax, float delta_z,
      float gamma_n, int B, int A, int Z, float* __restrict__ loss, float* __restrict__ grad, float* __restrict__ m_out,
      int64_t* __restrict__ astar_out) {
  __shared__ C51Scratch s_sc[C51_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * C51_WARPS + warp;
  if (i >= B) return;
  const int act = (int)actions[i];
  float g[C51_R];
  c51_core<C51_R>(s_sc[warp], lane, i, B, A, Z, q_on_ns + (size_t)i * A * Z, q_tg_ns + (size_t)i * A * Z,
           q_on_s + ((size_t)i * A + act) * Z, __ldg(returns + i), __ldg(nonterminals + i), __ldg(weights + i), support, vmin,
           vmax, delta_z, gamma_n, loss, m_out, astar_out, g);
  float* gq = grad + (size_t)i * A * Z;
  for (int j = lane; j < A * Z; j += 32) gq[j] = 0.0f;
  __syncwarp();
#pragma unroll
  for (int r = 0; r < C51_R; ++r) {
    int z = lane + 32 * r;
    if (z < Z) gq[(size_t)act * Z + z] = g[r];
  }
}

// Dueling entry point: fed by the fused heads' outputs z = (z_value | z_advantage) [rows][Z + A*Z]
// (online net: 2B rows, s then s'; target net: B rows).  ONE CTA PER SAMPLE, 8 warps: the kernel is pure dependent latency
// (184 KB in, 46 KB out), so the serial chain per sample is cut instead of packing samples into few CTAs:
//   phase 0  all threads stage the sample's three z rows in shared memory (every load in flight at once);
//   phase 1  warp a computes the expected value of action a of the online net on s' (model.py:75 dueling combination on the
//            fly, softmax, expectation) -- the A softmax rows of the double-DQN arg-max run side by side, not in sequence;
//   phase 2  warp 0 takes the arg-max (first maximum wins), assembles the two logit rows still needed (target net at a*,
//            online net at the taken action) and runs the projection / loss / gradient (c51_core, same arithmetic as the
//            plain entry point);
//   phase 3  all threads write dz:  dzv[z] = g[z],  dza[a][z] = g[z] * ([a == act] - 1/A).
constexpr int C51D_T = 256;

template <int C51_R>
__global__ void __launch_bounds__(C51D_T)
k_c51_dueling(const float* __restrict__ z_on, const float* __restrict__ z_tg, const int64_t* __restrict__ actions,
              const float* __restrict__ returns, const float* __restrict__ nonterminals, const float* __restrict__ weights,
              const float* __restrict__ support, float vmin, float vmax, float delta_z, float gamma_n, int B, int A, int Z,
              float* __restrict__ loss, float* __restrict__ dz, float* __restrict__ m_out, int64_t* __restrict__ astar_out) {
  extern __shared__ __align__(16) float s_dyn[];
  __shared__ C51Scratch s_sc;
  const int N2 = Z + A * Z;
  float* zs = s_dyn;              // [3][N2]: online(s), online(s'), target(s')
  float* q_t = zs + 3 * N2;       // [Z] target logits of a*
  float* q_s = q_t + Z;           // [Z] online logits of the taken action
  float* s_g = q_s + Z;           // [Z] gradient row
  float* s_ev = s_g + Z;          // [A] expected values
  const int i = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  {
    const int total = 3 * N2;
    const float* src[3] = {z_on + (size_t)i * N2, z_on + (size_t)(B + i) * N2, z_tg + (size_t)i * N2};
    for (int base = tid; base < total; base += C51D_T * 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {  // eight independent loads in flight per thread, then the stores
        const int idx = base + u * C51D_T;
        v[u] = 0.0f;
        if (idx < total) {
          const int t = idx / N2;
          v[u] = __ldg(src[t] + (idx - t * N2));
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int idx = base + u * C51D_T;
        if (idx < total) zs[idx] = v[u];
      }
    }
  }
  __syncthreads();
  const int act = (int)actions[i];
  float sup[C51_R];
#pragma unroll
  for (int r = 0; r < C51_R; ++r) sup[r] = (lane + 32 * r < Z) ? __ldg(support + lane + 32 * r) : 0.0f;
  {  // phase 1: expected value of every action of online(s'), one warp per action
    const float* r1 = zs + N2;
    float mean[C51_R];
#pragma unroll
    for (int r = 0; r < C51_R; ++r) {
      const int c = lane + 32 * r;
      float acc = 0.0f;
      if (c < Z)
        for (int a = 0; a < A; ++a) acc += r1[Z + a * Z + c];
      mean[r] = acc / (float)A;
    }
    for (int a = warp; a < A; a += C51D_T / 32) {
      float x[C51_R];
#pragma unroll
      for (int r = 0; r < C51_R; ++r) {
        const int c = lane + 32 * r;
        x[r] = (c < Z) ? r1[c] + r1[Z + a * Z + c] - mean[r] : -CUDART_INF_F;
      }
      const float ev = c51_expected_value<C51_R>(x, sup, Z, lane);
      if (lane == 0) s_ev[a] = ev;
    }
  }
  __syncthreads();
  if (warp == 0) {  // phase 2
    int best = 0;
    float best_ev = -CUDART_INF_F;
    for (int a = 0; a < A; ++a) {
      const float ev = s_ev[a];
      if (ev > best_ev) {  // first maximum wins, like torch.argmax
        best_ev = ev;
        best = a;
      }
    }
    for (int c = lane; c < Z; c += 32) {
      {  // online(s): the taken action only
        const float* r = zs;
        float mean = 0.0f;
        for (int a = 0; a < A; ++a) mean += r[Z + a * Z + c];
        q_s[c] = r[c] + r[Z + act * Z + c] - mean / (float)A;
      }
      {  // target(s') at a*
        const float* r = zs + 2 * N2;
        float mean = 0.0f;
        for (int a = 0; a < A; ++a) mean += r[Z + a * Z + c];
        mean = mean / (float)A;
        q_t[c] = r[c] + r[Z + best * Z + c] - mean;
      }
    }
    __syncwarp();
    float g[C51_R];
    c51_core<C51_R>(s_sc, lane, i, B, A, Z, nullptr, q_t, q_s, __ldg(returns + i), __ldg(nonterminals + i), __ldg(weights + i),
                    support, vmin, vmax, delta_z, gamma_n, loss, m_out, astar_out, g, best);
#pragma unroll
    for (int r = 0; r < C51_R; ++r)
      if (lane + 32 * r < Z) s_g[lane + 32 * r] = g[r];
  }
  __syncthreads();
  {  // phase 3
    float* dzi = dz + (size_t)i * N2;
    const float inv_a = 1.0f / (float)A;
    for (int idx = tid; idx < N2; idx += C51D_T) {
      if (idx < Z) {
        dzi[idx] = s_g[idx];
      } else {
        const int a = (idx - Z) / Z, c = (idx - Z) - a * Z;
        dzi[idx] = s_g[c] * ((a == act ? 1.0f : 0.0f) - inv_a);
      }
    }
  }
}

// ================================================================================================
// Q-values for acting / evaluation (agent.py:53-55 act, :110-112 evaluate_q): one warp per state.
// From the head output z = (z_value | z_advantage): q[a][z] = zv + za[a] - mean_a za (model.py:75), softmax over
// atoms (model.py:79), expected value sum_z support_z p_z (agent.py:55), then the arg-max / max over actions --
// everything after the network body in ONE launch, results stay on the device (no .item() per state).
// ================================================================================================
__global__ void __launch_bounds__(128)
k_q_select(const float* __restrict__ z, int M, int A, int Z, const float* __restrict__ support, float* __restrict__ q_out,
           int64_t* __restrict__ best_action, float* __restrict__ best_q) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * 4 + warp;
  if (m >= M) return;
  const float* zr = z + (size_t)m * (Z + A * Z);
  constexpr int R = RB_MAX_ATOMS / 32;
  float sup[R], zv[R], mean[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int c = lane + 32 * r;
    sup[r] = zv[r] = mean[r] = 0.0f;
    if (c < Z) {
      sup[r] = __ldg(support + c);
      zv[r] = __ldg(zr + c);
      float acc = 0.0f;
      for (int a = 0; a < A; ++a) acc += __ldg(zr + Z + a * Z + c);
      mean[r] = acc / (float)A;
    }
  }
  int best = 0;
  float best_ev = -CUDART_INF_F;
  for (int a = 0; a < A; ++a) {
    float x[R], mx = -CUDART_INF_F;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int c = lane + 32 * r;
      x[r] = (c < Z) ? zv[r] + __ldg(zr + Z + a * Z + c) - mean[r] : -CUDART_INF_F;
      mx = fmaxf(mx, x[r]);
    }
    mx = warp_max(mx);
    float se = 0.0f, sn = 0.0f;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float ee = (lane + 32 * r < Z) ? expf(x[r] - mx) : 0.0f;
      se = __fadd_rn(se, ee);
      sn = __fadd_rn(sn, __fmul_rn(sup[r], ee));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      se = __fadd_rn(se, __shfl_xor_sync(0xffffffffu, se, o));
      sn = __fadd_rn(sn, __shfl_xor_sync(0xffffffffu, sn, o));
    }
    const float ev = __fdiv_rn(sn, se);
    if (q_out && lane == 0) q_out[(size_t)m * A + a] = ev;
    if (ev > best_ev) {   // first maximum wins, like torch.argmax / max
      best_ev = ev;
      best = a;
    }
  }
  if (lane == 0) {
    if (best_action) best_action[m] = best;
    if (best_q) best_q[m] = best_ev;
  }
}

// ================================================================================================
// K6  noisy_resample : factorised Gaussian noise for every NoisyLinear of one net, one launch.
// ================================================================================================
struct NoisyPlan {
  float* w[RB_MAX_NOISY_LAYERS];
  float* b[RB_MAX_NOISY_LAYERS];
  int in_f[RB_MAX_NOISY_LAYERS];
  int out_f[RB_MAX_NOISY_LAYERS];
  int in_off[RB_MAX_NOISY_LAYERS];   // offset of the layer's eps_in in the concatenated normal stream
  int out_off[RB_MAX_NOISY_LAYERS];
  int rows_per_cta[RB_MAX_NOISY_LAYERS];
  int cta_begin[RB_MAX_NOISY_LAYERS + 1];
  int n;
};

constexpr int NOISY_THREADS = 256;

__global__ void __launch_bounds__(NOISY_THREADS)
k_noisy_resample(const __grid_constant__ NoisyPlan plan, const float* __restrict__ x_in, const float* __restrict__ x_out,
                 uint64_t seed, unsigned long long* rng_counter, int prescaled) {
  extern __shared__ __align__(16) float s_in[];  // f(eps_in) of this CTA's layer
  int l = 0;
  while (l + 1 < plan.n && (int)blockIdx.x >= plan.cta_begin[l + 1]) ++l;
  const int in_f = plan.in_f[l], out_f = plan.out_f[l];
  const int tile = blockIdx.x - plan.cta_begin[l];
  const int r0 = tile * plan.rows_per_cta[l], r1 = min(out_f, r0 + plan.rows_per_cta[l]);
  const unsigned long long ctr = rng_counter ? *rng_counter : 0ull;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (x_in) {
    for (int i = tid; i < in_f; i += NOISY_THREADS) {
      const float v = __ldg(x_in + plan.in_off[l] + i);
      s_in[i] = prescaled ? v : scale_noise(v);
    }
  } else {
    // global normal index g = in_off + i ; Philox block g/4 yields normals 4*(g/4) .. +3
    const int g0 = plan.in_off[l], g1 = g0 + in_f;
    for (int blk = g0 / 4 + tid; blk * 4 < g1; blk += NOISY_THREADS) {
      float4 z = normal4(seed, ctr, 0u, (uint32_t)blk);
      float zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        int g = blk * 4 + q;
        if (g >= g0 && g < g1) s_in[g - g0] = scale_noise(zz[q]);
      }
    }
  }
  __syncthreads();

  float* w = plan.w[l];
  float* bias = plan.b[l];
  const bool vec = (in_f % 4 == 0) && ((reinterpret_cast<uintptr_t>(w) & 15) == 0);
  for (int o = r0 + warp; o < r1; o += NOISY_THREADS / 32) {
    float xo;
    if (x_out) {
      xo = __ldg(x_out + plan.out_off[l] + o);
    } else {
      int g = plan.out_off[l] + o;
      float4 z = normal4(seed, ctr, 1u, (uint32_t)(g >> 2));
      int q = g & 3;
      xo = q == 0 ? z.x : (q == 1 ? z.y : (q == 2 ? z.z : z.w));
    }
    const float eo = (x_out && prescaled) ? xo : scale_noise(xo);
    if (lane == 0) bias[o] = eo;                       // model.py:40
    float* row = w + (size_t)o * in_f;                 // model.py:39 eps_out (outer) eps_in
    if (vec) {
      float4* row4 = reinterpret_cast<float4*>(row);
      const float4* in4 = reinterpret_cast<const float4*>(s_in);
      for (int c = lane; c < in_f / 4; c += 32) {
        float4 e = in4[c];
        __stcs(row4 + c, make_float4(__fmul_rn(eo, e.x), __fmul_rn(eo, e.y), __fmul_rn(eo, e.z), __fmul_rn(eo, e.w)));
      }
    } else {
      for (int c = lane; c < in_f; c += 32) row[c] = __fmul_rn(eo, s_in[c]);
    }
  }
}

// counter bump runs as its own tiny kernel after the resample grid (all CTAs must read the old value)
__global__ void k_bump_counter(unsigned long long* ctr) { *ctr += 1ull; }

// model.py:43-44: W = mu + sigma * eps
__global__ void __launch_bounds__(256)
k_noisy_compose(const float* __restrict__ mu, const float* __restrict__ sigma, const float* __restrict__ eps,
                int64_t count, float* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if ((count & 3) == 0 && (((uintptr_t)mu | (uintptr_t)sigma | (uintptr_t)eps | (uintptr_t)out) & 15) == 0) {
    const int64_t n4 = count >> 2;
    for (; i < n4; i += stride) {
      float4 m = __ldg(reinterpret_cast<const float4*>(mu) + i), s = __ldg(reinterpret_cast<const float4*>(sigma) + i),
             e = __ldg(reinterpret_cast<const float4*>(eps) + i);
      reinterpret_cast<float4*>(out)[i] =
          make_float4(__fadd_rn(m.x, __fmul_rn(s.x, e.x)), __fadd_rn(m.y, __fmul_rn(s.y, e.y)),
                      __fadd_rn(m.z, __fmul_rn(s.z, e.z)), __fadd_rn(m.w, __fmul_rn(s.w, e.w)));
    }
  } else {
    for (; i < count; i += stride) out[i] = __fadd_rn(mu[i], __fmul_rn(sigma[i], eps[i]));
  }
}

// ================================================================================================
// K7  clip_adam : global-norm clip + Adam on flat buffers (agent.py:97-98), two launches.
// ================================================================================================
constexpr int ADAM_THREADS = 256;
constexpr int ADAM_MAX_CTAS = 148 * 8;

__global__ void __launch_bounds__(ADAM_THREADS)
k_sqnorm(const float* __restrict__ grad, int64_t P, float grad_scale, double* __restrict__ partial) {
  __shared__ double s_red[ADAM_THREADS / 32];
  double acc = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if ((P & 3) == 0 && ((uintptr_t)grad & 15) == 0) {
    for (; i < (P >> 2); i += stride) {
      float4 g = __ldg(reinterpret_cast<const float4*>(grad) + i);
      float a = g.x * grad_scale, b = g.y * grad_scale, c = g.z * grad_scale, d = g.w * grad_scale;
      acc += (double)a * a + (double)b * b + (double)c * c + (double)d * d;
    }
  } else {
    for (; i < P; i += stride) {
      float a = grad[i] * grad_scale;
      acc += (double)a * a;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < ADAM_THREADS / 32; ++w) t += s_red[w];
    partial[blockIdx.x] = t;
  }
}

// sqrt / reciprocal through the SFU approximations (<= 2 ulp): the kernel was partly instruction bound on the
// IEEE div/sqrt sequences (r01 ncu: SM throughput 55 %), and the update tolerates 1e-6 relative error.
__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float coef, float b1, float b2,
                                         float step_size, float inv_bc2_sqrt, float eps) {
  g = g * coef;
  m = fmaf(g - m, 1.0f - b1, m);                 // exp_avg.lerp_(grad, 1 - beta1)
  v = fmaf(v, b2, (1.0f - b2) * g * g);          // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = fmaf(fast_sqrt(v), inv_bc2_sqrt, eps);
  p = fmaf(-step_size * m, fast_rcp(denom), p);  // param.addcdiv_(exp_avg, denom, value=-step_size)
}

__global__ void __launch_bounds__(ADAM_THREADS)
k_clip_adam(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ exp_avg,
            float* __restrict__ exp_avg_sq, int64_t P, float grad_scale, float max_norm, float lr, float b1, float b2,
            float eps, int64_t* __restrict__ step_count, const double* __restrict__ partial, int n_partial,
            float* __restrict__ norm_out, unsigned int* __restrict__ done_ticket, const int32_t* __restrict__ gate) {
  const bool skip = gate && *gate == 0;   // rejected sample batch: no parameter update, no step (see k_tree_sample)
  __shared__ double s_red[ADAM_THREADS / 32];
  __shared__ float s_coef;
  // every CTA re-reduces the (few hundred) partial sums in the same order: deterministic, no atomics
  double acc = 0.0;
  for (int i = threadIdx.x; i < n_partial; i += ADAM_THREADS) acc += partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < ADAM_THREADS / 32; ++w) t += s_red[w];
    float norm = (float)sqrt(t);
    float c = max_norm / (norm + 1e-6f);         // torch clip_grad_norm_
    s_coef = fminf(c, 1.0f) * grad_scale;
    if (norm_out && blockIdx.x == 0) *norm_out = norm;
  }
  __syncthreads();
  const float coef = s_coef;
  const int64_t step = *step_count + 1;
  const double bc1 = 1.0 - pow((double)b1, (double)step);
  const double bc2 = 1.0 - pow((double)b2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float bc2_sqrt = (float)(1.0 / sqrt(bc2));  // passed to adam_one as the reciprocal

  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool vec = (P & 3) == 0 && (((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0;
  if (skip) {
  } else if (vec) {
    for (; i < (P >> 2); i += stride) {
      float4 p = reinterpret_cast<float4*>(param)[i], g = __ldg(reinterpret_cast<const float4*>(grad) + i),
             m = reinterpret_cast<float4*>(exp_avg)[i], v = reinterpret_cast<float4*>(exp_avg_sq)[i];
      adam_one(p.x, g.x, m.x, v.x, coef, b1, b2, step_size, bc2_sqrt, eps);
      adam_one(p.y, g.y, m.y, v.y, coef, b1, b2, step_size, bc2_sqrt, eps);
      adam_one(p.z, g.z, m.z, v.z, coef, b1, b2, step_size, bc2_sqrt, eps);
      adam_one(p.w, g.w, m.w, v.w, coef, b1, b2, step_size, bc2_sqrt, eps);
      reinterpret_cast<float4*>(param)[i] = p;
      reinterpret_cast<float4*>(exp_avg)[i] = m;
      reinterpret_cast<float4*>(exp_avg_sq)[i] = v;
    }
  } else {
    for (; i < P; i += stride) {
      float p = param[i], m = exp_avg[i], v = exp_avg_sq[i];
      adam_one(p, grad[i], m, v, coef, b1, b2, step_size, bc2_sqrt, eps);
      param[i] = p;
      exp_avg[i] = m;
      exp_avg_sq[i] = v;
    }
  }
  // every CTA has read *step_count above; the last one to get here advances it (self-resetting ticket), which saves a
  // dependent 1-thread launch on the critical path
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(done_ticket, 1u);
    if (t == gridDim.x - 1) {
      *done_ticket = 0u;
      if (!skip) *step_count = step;
    }
  }
}

int adam_ctas(int64_t P) {
  int64_t want = (P / 4 + ADAM_THREADS - 1) / ADAM_THREADS;
  if (want < 1) want = 1;
  if (want > ADAM_MAX_CTAS) want = ADAM_MAX_CTAS;
  return (int)want;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int rb_abi_version(void) { return RB_ABI_VERSION; }

int rb_profile_enable(int on) {
  rbi::g_prof_on = on != 0;
  return RB_OK;
}

int rb_profile_collect(int kernel_id, double* total_ms, int* launches) {
  if (kernel_id < 0 || kernel_id >= RB_KERNEL_COUNT || !total_ms || !launches) return fail(RB_ERR_INVAL, "rb_profile_collect: bad argument");
  ProfKernel* k = &rbi::g_prof[kernel_id];
  double t = 0.0;
  for (int i = 0; i < k->used; ++i) {
    float ms = 0.0f;
    cudaError_t e = cudaEventSynchronize(k->e1[i]);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, k->e0[i], k->e1[i]);
    if (e != cudaSuccess) return fail(RB_ERR_CUDA, cudaGetErrorString(e));
    t += ms;
  }
  *total_ms = t;
  *launches = k->used;
  k->used = 0;
  return RB_OK;
}

const char* rb_last_error(void) { return rbi::g_err; }

int rb_tree_update(float* tree, int64_t tree_start, int64_t size, const int64_t* tree_idx, const float* raw_priority,
                   float omega, int omega_is_applied, int B, float* running_max, int32_t* status, const int32_t* gate,
                   rb_stream_t stream) {
  if (!tree || !tree_idx || !raw_priority || !running_max) return fail(RB_ERR_INVAL, "rb_tree_update: null pointer");
  if (B <= 0 || size <= 0 || (size & 1)) return fail(RB_ERR_INVAL, "rb_tree_update: B > 0 and an even size are required");
  if (tree_start + size > ((int64_t)1 << 31)) return fail(RB_ERR_RANGE, "rb_tree_update: tree larger than 2^31 nodes");
  { ProfScope prof_(RB_K_TREE_UPDATE, (cudaStream_t)stream);
    if (B <= 32 && tree_depth(tree_start) <= 30)
      k_tree_update_warp<<<1, 32, 0, (cudaStream_t)stream>>>(tree, tree_start, size, tree_idx, raw_priority, omega,
                                                            omega_is_applied, B, running_max, status, gate);
    else
      k_tree_update<<<1, UPD_THREADS, 0, (cudaStream_t)stream>>>(tree, tree_start, size, tree_idx, raw_priority, omega,
                                                               omega_is_applied, B, running_max, status, gate); }
  return check_launch("rb_tree_update");
}

int rb_tree_find(const float* tree, int64_t tree_start, int64_t size, const double* values, int B, float* probs,
                 int64_t* data_idx, int64_t* tree_idx, rb_stream_t stream) {
  if (!tree || !values || !probs || !data_idx || !tree_idx) return fail(RB_ERR_INVAL, "rb_tree_find: null pointer");
  if (B <= 0 || size <= 0) return fail(RB_ERR_INVAL, "rb_tree_find: B and size must be positive");
  int ctas = (B + 31) / 32;
  if (ctas > 148) ctas = 148;
  { ProfScope prof_(RB_K_TREE_FIND, (cudaStream_t)stream);
    k_tree_find<<<ctas, SAMPLE_THREADS, 0, (cudaStream_t)stream>>>(tree, tree_start, size, values, B, probs, data_idx, tree_idx); }
  return check_launch("rb_tree_find");
}

int rb_tree_sample(const float* tree, int64_t tree_start, int64_t size, const int64_t* ring_state, int n, int history,
                   const double* u01, int u01_attempts, uint64_t seed, uint64_t* rng_counter, int B, float beta,
                   const float* beta_dev, int max_attempts, float* probs, int64_t* data_idx, int64_t* tree_idx,
                   float* weights, int32_t* status, rb_stream_t stream) {
  if (!tree || !ring_state || !probs || !data_idx || !tree_idx || !weights || !status)
    return fail(RB_ERR_INVAL, "rb_tree_sample: null pointer");
  if (B <= 0 || size <= 0 || (size & 1)) return fail(RB_ERR_INVAL, "rb_tree_sample: B > 0 and an even size are required");
  if (u01 == nullptr && rng_counter == nullptr) return fail(RB_ERR_INVAL, "rb_tree_sample: need u01 or rng_counter");
  if ((u01 != nullptr && u01_attempts <= 0) || (u01 == nullptr && max_attempts <= 0))
    return fail(RB_ERR_INVAL, "rb_tree_sample: attempts must be positive");
  { ProfScope prof_(RB_K_TREE_SAMPLE, (cudaStream_t)stream);
    k_tree_sample<<<1, SAMPLE_THREADS, 0, (cudaStream_t)stream>>>(
      tree, tree_start, size, ring_state, n, history, u01, u01_attempts, seed, (unsigned long long*)rng_counter, B, beta,
      beta_dev, max_attempts, probs, data_idx, tree_idx, weights, status); }
  return check_launch("rb_tree_sample");
}

static int gather_split(int ctas_without_split) {
  // aim for >= 4 CTAs per SM (148 SMs) so small batches still cover the machine; a frame is 441 x 16 B
  int split = 1;
  while (split < 2 && ctas_without_split * split < 148 * 2) split *= 2;  // 221 of 256 threads busy per CTA at split 2
  return split;
}

int rb_gather(const uint8_t* frames, const int32_t* timestep, const int32_t* action, const float* reward,
              const uint8_t* nonterminal, int64_t size, const int64_t* data_idx, int B, int history, int n,
              const float* gamma_pow, float* states, float* next_states, int64_t* actions, float* returns,
              float* nonterminals, rb_stream_t stream) {
  if (!frames || !timestep || !action || !reward || !nonterminal || !data_idx || !gamma_pow || !states || !next_states ||
      !actions || !returns || !nonterminals)
    return fail(RB_ERR_INVAL, "rb_gather: null pointer");
  if (B <= 0 || history <= 0 || n <= 0 || size <= 0) return fail(RB_ERR_INVAL, "rb_gather: sizes must be positive");
  if (history + n > RB_MAX_WINDOW) return fail(RB_ERR_RANGE, "rb_gather: history + n exceeds RB_MAX_WINDOW");
  if (B > 65535) return fail(RB_ERR_RANGE, "rb_gather: B exceeds 65535");
  const int used = (history + n < 2 * history) ? history + n : 2 * history;
  const int split = gather_split(used * B);
  dim3 grid(used * split, B);
  { ProfScope prof_(RB_K_GATHER, (cudaStream_t)stream);
    k_gather<<<grid, GATHER_THREADS, 0, (cudaStream_t)stream>>>(frames, timestep, action, reward, nonterminal, size, data_idx,
                                                              B, history, n, gamma_pow, states, next_states, actions,
                                                              returns, nonterminals, split); }
  return check_launch("rb_gather");
}

int rb_iter_states(const uint8_t* frames, const int32_t* timestep, int64_t size, int64_t first, int count, int history,
                   float* out, rb_stream_t stream) {
  if (!frames || !timestep || !out) return fail(RB_ERR_INVAL, "rb_iter_states: null pointer");
  if (count <= 0 || history <= 0 || history > RB_MAX_WINDOW || count > 65535)
    return fail(RB_ERR_RANGE, "rb_iter_states: count/history out of range");
  dim3 grid(history, count);
  { ProfScope prof_(RB_K_ITER_STATES, (cudaStream_t)stream);
    k_iter_states<<<grid, GATHER_THREADS, 0, (cudaStream_t)stream>>>(frames, timestep, size, first, history, out); }
  return check_launch("rb_iter_states");
}

int rb_append(float* tree, int64_t tree_start, int64_t size, uint8_t* frames, int32_t* timestep, int32_t* action,
              float* reward, uint8_t* nonterminal, int64_t* ring_state, float* running_max, const float* state_last_frame,
              int32_t action_value, float reward_value, int terminal, rb_stream_t stream) {
  if (!tree || !frames || !timestep || !action || !reward || !nonterminal || !ring_state || !running_max || !state_last_frame)
    return fail(RB_ERR_INVAL, "rb_append: null pointer");
  if (size <= 0 || (size & 1)) return fail(RB_ERR_INVAL, "rb_append: an even size is required");
  if (tree_depth(tree_start) > 32) return fail(RB_ERR_RANGE, "rb_append: tree deeper than 32 levels");  // one lane per level
  if (((uintptr_t)state_last_frame & 15) != 0) return fail(RB_ERR_INVAL, "rb_append: state_last_frame must be 16-byte aligned");
  { ProfScope prof_(RB_K_APPEND, (cudaStream_t)stream);
    k_append<<<1, APPEND_THREADS, 0, (cudaStream_t)stream>>>(tree, tree_start, size, frames, timestep, action, reward,
                                                           nonterminal, ring_state, running_max, state_last_frame,
                                                           action_value, reward_value, terminal); }
  return check_launch("rb_append");
}

int rb_append_batch(float* tree, int64_t tree_start, int64_t size, uint8_t* frames, int32_t* timestep, int32_t* action,
                    float* reward, uint8_t* nonterminal, int64_t* ring_state, float* running_max,
                    const float* const* last_frames, const int32_t* actions, const float* rewards, const int32_t* terminals,
                    int k, rb_stream_t stream) {
  if (!tree || !frames || !timestep || !action || !reward || !nonterminal || !ring_state || !running_max || !last_frames ||
      !actions || !rewards || !terminals)
    return fail(RB_ERR_INVAL, "rb_append_batch: null pointer");
  if (size <= 0 || (size & 1)) return fail(RB_ERR_INVAL, "rb_append_batch: an even size is required");
  if (k <= 0 || k > RB_APPEND_BATCH || k > size) return fail(RB_ERR_RANGE, "rb_append_batch: 1 <= k <= RB_APPEND_BATCH (and k <= size)");
  if (tree_depth(tree_start) > 30) return fail(RB_ERR_RANGE, "rb_append_batch: tree deeper than 30 levels");
  AppendBatch ab;
  memset(&ab, 0, sizeof(ab));
  ab.k = k;
  for (int j = 0; j < k; ++j) {
    if (!last_frames[j] || ((uintptr_t)last_frames[j] & 15)) return fail(RB_ERR_INVAL, "rb_append_batch: frames must be non-null and 16-byte aligned");
    ab.frame[j] = last_frames[j];
    ab.action[j] = actions[j];
    ab.reward[j] = rewards[j];
    ab.terminal[j] = terminals[j] ? 1 : 0;
  }
  { ProfScope prof_(RB_K_APPEND, (cudaStream_t)stream);
    k_append_batch<<<k, APPEND_THREADS, 0, (cudaStream_t)stream>>>(tree, tree_start, size, frames, timestep, action, reward,
                                                                   nonterminal, ring_state, running_max, ab); }
  return check_launch("rb_append_batch");
}

int rb_c51_loss_grad(const float* q_online_s, const float* q_online_ns, const float* q_target_ns, const int64_t* actions,
                     const float* returns, const float* nonterminals, const float* weights, const float* support,
                     float vmin, float vmax, float delta_z, float gamma_n, int B, int A, int Z, float* loss,
                     float* grad_q_online_s, float* m_out, int64_t* astar_out, rb_stream_t stream) {
  if (!q_online_s || !q_online_ns || !q_target_ns || !actions || !returns || !nonterminals || !weights || !support ||
      !loss || !grad_q_online_s)
    return fail(RB_ERR_INVAL, "rb_c51_loss_grad: null pointer");
  if (B <= 0 || A <= 0 || Z <= 1) return fail(RB_ERR_INVAL, "rb_c51_loss_grad: B, A > 0 and Z > 1 are required");
  if (Z > RB_MAX_ATOMS) return fail(RB_ERR_RANGE, "rb_c51_loss_grad: Z exceeds RB_MAX_ATOMS");
  const int ctas = (B + C51_WARPS - 1) / C51_WARPS;
  { ProfScope prof_(RB_K_C51, (cudaStream_t)stream);
    if (Z <= 64)
      k_c51<2><<<ctas, C51_WARPS * 32, 0, (cudaStream_t)stream>>>(q_online_s, q_online_ns, q_target_ns, actions, returns,
                                                               nonterminals, weights, support, vmin, vmax, delta_z, gamma_n,
                                                               B, A, Z, loss, grad_q_online_s, m_out, astar_out);
    else
      k_c51<4><<<ctas, C51_WARPS * 32, 0, (cudaStream_t)stream>>>(q_online_s, q_online_ns, q_target_ns, actions, returns,
                                                               nonterminals, weights, support, vmin, vmax, delta_z, gamma_n,
                                                               B, A, Z, loss, grad_q_online_s, m_out, astar_out); }
  return check_launch("rb_c51_loss_grad");
}

static int noisy_launch(float* const* weight_eps, float* const* bias_eps, const int* in_features, const int* out_features,
                        int n_layers, const float* x_in, const float* x_out, uint64_t seed, uint64_t* rng_counter,
                        int prescaled, rb_stream_t stream) {
  if (!weight_eps || !bias_eps || !in_features || !out_features) return fail(RB_ERR_INVAL, "rb_noisy_resample: null pointer");
  if (n_layers <= 0 || n_layers > RB_MAX_NOISY_LAYERS) return fail(RB_ERR_RANGE, "rb_noisy_resample: n_layers out of range");
  if ((x_in == nullptr) != (x_out == nullptr)) return fail(RB_ERR_INVAL, "rb_noisy_resample: give both x_in and x_out or neither");
  if (x_in == nullptr && rng_counter == nullptr) return fail(RB_ERR_INVAL, "rb_noisy_resample: need injected normals or rng_counter");
  NoisyPlan plan;
  memset(&plan, 0, sizeof(plan));
  plan.n = n_layers;
  int in_off = 0, out_off = 0, ctas = 0, max_in = 0;
  for (int l = 0; l < n_layers; ++l) {
    if (!weight_eps[l] || !bias_eps[l] || in_features[l] <= 0 || out_features[l] <= 0)
      return fail(RB_ERR_INVAL, "rb_noisy_resample: bad layer description");
    if (in_features[l] > 57000) return fail(RB_ERR_RANGE, "rb_noisy_resample: in_features exceeds shared-memory staging");
    plan.w[l] = weight_eps[l];
    plan.b[l] = bias_eps[l];
    plan.in_f[l] = in_features[l];
    plan.out_f[l] = out_features[l];
    plan.in_off[l] = in_off;
    plan.out_off[l] = out_off;
    in_off += in_features[l];
    out_off += out_features[l];
    // ~24K weights per CTA, at least 8 rows (one per warp)
    int rows = 24576 / in_features[l];
    if (rows < 8) rows = 8;
    plan.rows_per_cta[l] = rows;
    plan.cta_begin[l] = ctas;
    ctas += (out_features[l] + rows - 1) / rows;
    if (in_features[l] > max_in) max_in = in_features[l];
  }
  plan.cta_begin[n_layers] = ctas;
  const size_t smem = (size_t)max_in * sizeof(float);
  {
    int rc_s = rbi::ensure_dynamic_smem(k_noisy_resample, smem, "rb_noisy_resample");
    if (rc_s != RB_OK) return rc_s;
  }
  { ProfScope prof_(RB_K_NOISY_RESAMPLE, (cudaStream_t)stream);
    k_noisy_resample<<<ctas, NOISY_THREADS, smem, (cudaStream_t)stream>>>(plan, x_in, x_out, seed,
                                                                       (unsigned long long*)rng_counter, prescaled); }
  int rc = check_launch("rb_noisy_resample");
  if (rc != RB_OK) return rc;
  if (x_in == nullptr) {
    k_bump_counter<<<1, 1, 0, (cudaStream_t)stream>>>((unsigned long long*)rng_counter);
    rc = check_launch("rb_noisy_resample(counter)");
  }
  return rc;
}

int rb_noisy_resample(float* const* weight_eps, float* const* bias_eps, const int* in_features, const int* out_features,
                      int n_layers, const float* x_in, const float* x_out, uint64_t seed, uint64_t* rng_counter,
                      rb_stream_t stream) {
  return noisy_launch(weight_eps, bias_eps, in_features, out_features, n_layers, x_in, x_out, seed, rng_counter, 0, stream);
}

int rb_noisy_outer(float* const* weight_eps, float* const* bias_eps, const int* in_features, const int* out_features,
                   int n_layers, const float* f_in, const float* f_out, rb_stream_t stream) {
  if (!f_in || !f_out) return fail(RB_ERR_INVAL, "rb_noisy_outer: null factor vectors");
  return noisy_launch(weight_eps, bias_eps, in_features, out_features, n_layers, f_in, f_out, 0, nullptr, 1, stream);
}

int rb_c51_dueling_loss_grad(const float* z_online, const float* z_target, int actions_n, int atoms, const int64_t* actions,
                             const float* returns, const float* nonterminals, const float* weights, const float* support,
                             float vmin, float vmax, float delta_z, float gamma_n, int B, float* loss, float* dz, float* m_out,
                             int64_t* astar_out, rb_stream_t stream) {
  if (!z_online || !z_target || !actions || !returns || !nonterminals || !weights || !support || !loss || !dz)
    return fail(RB_ERR_INVAL, "rb_c51_dueling_loss_grad: null pointer");
  const int Z = atoms, A = actions_n;
  if (B <= 0 || A <= 0 || Z <= 1) return fail(RB_ERR_INVAL, "rb_c51_dueling_loss_grad: B, actions > 0 and atoms > 1 are required");
  if (Z > RB_MAX_ATOMS) return fail(RB_ERR_RANGE, "rb_c51_dueling_loss_grad: atoms exceeds RB_MAX_ATOMS");
  const size_t smem = (size_t)(3 * (Z + A * Z) + 3 * Z + A) * sizeof(float);
  if (smem > 200 * 1024) return fail(RB_ERR_RANGE, "rb_c51_dueling_loss_grad: actions * atoms too large");
  int rc_s = rbi::ensure_dynamic_smem(k_c51_dueling<2>, smem, "rb_c51_dueling_loss_grad");
  if (rc_s == RB_OK) rc_s = rbi::ensure_dynamic_smem(k_c51_dueling<4>, smem, "rb_c51_dueling_loss_grad");
  if (rc_s != RB_OK) return rc_s;
  { ProfScope prof_(RB_K_C51_DUELING, (cudaStream_t)stream);
    if (Z <= 64)
      k_c51_dueling<2><<<B, C51D_T, smem, (cudaStream_t)stream>>>(z_online, z_target, actions, returns, nonterminals, weights,
                                                                  support, vmin, vmax, delta_z, gamma_n, B, A, Z, loss, dz, m_out,
                                                                  astar_out);
    else
      k_c51_dueling<4><<<B, C51D_T, smem, (cudaStream_t)stream>>>(z_online, z_target, actions, returns, nonterminals, weights,
                                                                  support, vmin, vmax, delta_z, gamma_n, B, A, Z, loss, dz, m_out,
                                                                  astar_out); }
  return check_launch("rb_c51_dueling_loss_grad");
}

int rb_q_values(const float* z, int M, int actions, int atoms, const float* support, float* q, int64_t* best_action,
                float* best_q, rb_stream_t stream) {
  if (!z || !support) return fail(RB_ERR_INVAL, "rb_q_values: null pointer");
  if (!q && !best_action && !best_q) return fail(RB_ERR_INVAL, "rb_q_values: no output requested");
  if (M <= 0 || actions <= 0 || atoms <= 1) return fail(RB_ERR_INVAL, "rb_q_values: M, actions > 0 and atoms > 1 are required");
  if (atoms > RB_MAX_ATOMS) return fail(RB_ERR_RANGE, "rb_q_values: atoms exceeds RB_MAX_ATOMS");
  { ProfScope prof_(RB_K_Q_VALUES, (cudaStream_t)stream);
    k_q_select<<<(M + 3) / 4, 128, 0, (cudaStream_t)stream>>>(z, M, actions, atoms, support, q, best_action, best_q); }
  return check_launch("rb_q_values");
}

int rb_noisy_compose(const float* mu, const float* sigma, const float* eps, int64_t count, float* out, rb_stream_t stream) {
  if (!mu || !sigma || !eps || !out) return fail(RB_ERR_INVAL, "rb_noisy_compose: null pointer");
  if (count <= 0) return fail(RB_ERR_INVAL, "rb_noisy_compose: count must be positive");
  int64_t ctas = (count / 4 + 255) / 256;
  if (ctas < 1) ctas = 1;
  if (ctas > 148 * 8) ctas = 148 * 8;
  { ProfScope prof_(RB_K_NOISY_COMPOSE, (cudaStream_t)stream);
    k_noisy_compose<<<(int)ctas, 256, 0, (cudaStream_t)stream>>>(mu, sigma, eps, count, out); }
  return check_launch("rb_noisy_compose");
}

int rb_clip_adam_scratch_elems(void) { return ADAM_MAX_CTAS + 1; }  // partial sums + one ticket word (zero-initialised by the caller)

int rb_clip_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t P, float grad_scale,
                 float max_norm, float lr, float beta1, float beta2, float eps, int64_t* step_count, double* partial_sums,
                 float* norm_out, const int32_t* gate, rb_stream_t stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || !step_count || !partial_sums)
    return fail(RB_ERR_INVAL, "rb_clip_adam: null pointer");
  if (P <= 0) return fail(RB_ERR_INVAL, "rb_clip_adam: P must be positive");
  const int ctas = adam_ctas(P);
  { ProfScope prof_(RB_K_SQNORM, (cudaStream_t)stream);
    k_sqnorm<<<ctas, ADAM_THREADS, 0, (cudaStream_t)stream>>>(grad, P, grad_scale, partial_sums); }
  int rc = check_launch("rb_clip_adam(norm)");
  if (rc != RB_OK) return rc;
  { ProfScope prof_(RB_K_CLIP_ADAM, (cudaStream_t)stream);
    k_clip_adam<<<ctas, ADAM_THREADS, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, P, grad_scale, max_norm, lr,
                                                               beta1, beta2, eps, step_count, partial_sums, ctas, norm_out,
                                                               reinterpret_cast<unsigned int*>(partial_sums + ADAM_MAX_CTAS), gate); }
  return check_launch("rb_clip_adam");
}

}  // extern "C"
