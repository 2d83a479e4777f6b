// (c) Aggregation epilogue: probability-weighted sum of the K cell outputs + gated skip.
// Reference: models/DynamicInteraction.py:50-67 (= :118-132, non-final) and :104-117 (final).
//
// HBM-bound.  Fusions that keep per-cell intermediates out of HBM:
//   * relu() of the RIC cell (Cells.py:36-40) is applied on the fly to the raw layer input;
//   * GLAC / GESC outputs are [B,D] vectors broadcast over L (Cells.py:173,209): never expanded;
//   * the mean over L of every output (the next layer's router input, Router.py:23) is produced
//     in the same pass (forward) and its gradient consumed in the same pass (backward);
//   * the gated skip of the final layer (:108-111) reads / writes a cell input only for samples whose
//     gate is actually set (p_j < 1e-4/K: essentially never), see d2r_gate_skip_bwd.
//
// Layout (round 2): every global access is 16 bytes -- a lane owns V = 16/sizeof(T) consecutive columns (8 bf16 or
// 4 fp32), a warp 32*V columns (512 bytes per row), a block of 4 warps owns those columns of ONE sample and walks
// all L rows, so the reductions over L (pooled means, broadcast-cell gradients, dP) need no second pass.  Per-thread
// state is kept small -- the pooled means are rebuilt from the column sums of the 4 full cells instead of 6
// per-output accumulators -- so that 4 blocks are resident per SM (forward) with 8 x 16 B loads in flight per
// thread: round 1's 8-byte / 168-register version had 24 KB in flight per SM and stopped at 0.42 of the HBM peak.
// The set of broadcast cells is a compile-time mask (cells 1 and 5 of emb_lst).
#include "common.cuh"
#include "ptx.cuh"

namespace d2r {
namespace {

constexpr int kWarps = 4;
constexpr int kThreads = kWarps * 32;

template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int V = 4;
  using Raw = float4;
  static __device__ __forceinline__ void unpack(const Raw& r, float (&v)[4]) { v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w; }
  static __device__ __forceinline__ Raw pack(const float (&v)[4]) { return make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct Vec<__nv_bfloat16> {
  static constexpr int V = 8;
  using Raw = uint4;
  static __device__ __forceinline__ void unpack(const Raw& r, float (&v)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x;
      v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ Raw pack(const float (&v)[8]) {
    Raw r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    return r;
  }
};
template <typename T> __device__ __forceinline__ typename Vec<T>::Raw ldraw(const T* p) {
  return *reinterpret_cast<const typename Vec<T>::Raw*>(p);
}
// outputs are written once and next read by a GEMM of the following layer, after several hundred MB of other
// traffic: streaming (evict-first) stores keep them from displacing the operands still being read
template <typename T> __device__ __forceinline__ void straw(T* p, const typename Vec<T>::Raw& r) {
  __stcs(reinterpret_cast<typename Vec<T>::Raw*>(p), r);
}
template <int V> __device__ __forceinline__ void ldf(const float* p, float (&v)[V]) {
#pragma unroll
  for (int q = 0; q < V; q += 4) {
    const float4 a = *reinterpret_cast<const float4*>(p + q);
    v[q] = a.x; v[q + 1] = a.y; v[q + 2] = a.z; v[q + 3] = a.w;
  }
}
template <int V> __device__ __forceinline__ void stf(float* p, const float (&v)[V]) {
#pragma unroll
  for (int q = 0; q < V; q += 4) *reinterpret_cast<float4*>(p + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
}

// broadcast cells of emb_lst: GLAC (1) and, with six cells, GESC (5)
template <int K> __host__ __device__ constexpr bool is_bcast(int j) { return j == 1 || (K == 6 && j == 5); }
// dense index of full cell j among the full cells (K=6: cells 0,2,3,4 -> 0,1,2,3; K=4: 0,2,3 -> 0,1,2)
template <int K> __host__ __device__ constexpr int full_idx(int j) { return j == 0 ? 0 : j - 1; }
template <int K> constexpr int num_full() { return K == 6 ? 4 : 3; }
// dense index of broadcast cell j among the broadcast cells (1 -> 0, 5 -> 1)
template <int K> __host__ __device__ constexpr int bc_idx(int j) { return j == 1 ? 0 : 1; }

struct AggP {
  int K, n_out;
  long long B, L, D;
  d2r_ptr8 full, bvec, inputs, out;
  const float* P;
  const float* gate;
  float* pooled;
  // backward
  d2r_ptr8 d_out, d_full, d_bvec;
  const float* d_pooled;
  float* dP;
};

// cross-warp sum of per-thread partials v[V] for this block's columns; red: [kWarps][32*V] floats.
// Every thread ends up with the block total of its own columns.
template <int V>
__device__ __forceinline__ void block_colsum(float (&v)[V], float* red, int warp, int lane) {
  constexpr int CB = 32 * V;
  __syncthreads();
#pragma unroll
  for (int q = 0; q < V; ++q) red[warp * CB + lane * V + q] = v[q];
  __syncthreads();
#pragma unroll
  for (int q = 0; q < V; ++q) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += red[w * CB + lane * V + q];
    v[q] = s;
  }
}

// ------------------------------------------------------------------ forward, non-final
template <typename T, int K>
__global__ void __launch_bounds__(kThreads, 3) agg_fwd_kernel(const AggP p) {
  constexpr int V = Vec<T>::V, CB = 32 * V, NF = num_full<K>();
  using Raw = typename Vec<T>::Raw;
  __shared__ float sP[K * K];
  __shared__ float sG[K];
  __shared__ __align__(16) float sBC[K * CB];     // per output: contribution of the broadcast cells (constant over L)
  __shared__ __align__(16) float red[kWarps * CB];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = blockIdx.y;
  const long long col = (long long)blockIdx.x * CB + lane * V;
  if (threadIdx.x < K * K) sP[threadIdx.x] = p.P[b * K * K + threadIdx.x];
  if (threadIdx.x < K) sG[threadIdx.x] = p.gate[b * K + threadIdx.x];
  __syncthreads();
  for (int idx = threadIdx.x; idx < K * CB; idx += kThreads) {
    const int i = idx / CB, c = idx % CB;
    const long long gc = (long long)blockIdx.x * CB + c;
    float a = 0.f;
    if (gc < p.D) {
#pragma unroll
      for (int j = 0; j < K; ++j)
        if (is_bcast<K>(j)) a = fmaf(sP[i * K + j], static_cast<const float*>(p.bvec.p[j])[b * p.D + gc], a);
    }
    sBC[idx] = a;
  }
  __syncthreads();
  const bool active = col < p.D;
  float cs[NF][V];                // column sums over this thread's rows of the full cells (relu applied to cell 0)
#pragma unroll
  for (int j = 0; j < NF; ++j)
#pragma unroll
    for (int q = 0; q < V; ++q) cs[j][q] = 0.f;
  if (active) {
    const long long base = b * p.L * p.D + col;
    // two rows per iteration: 2 * NF 16-byte loads are issued before the first is consumed
    for (long long l = warp; l < p.L; l += 2 * kWarps) {
      const bool two = l + kWarps < p.L;
      Raw r[2][NF];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (h == 1 && !two) break;
        const long long off = base + (l + h * kWarps) * p.D;
#pragma unroll
        for (int j = 0; j < K; ++j)
          if (!is_bcast<K>(j)) r[h][full_idx<K>(j)] = ldraw<T>(static_cast<const T*>(p.full.p[j]) + off);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (h == 1 && !two) break;
        const long long off = base + (l + h * kWarps) * p.D;
        float e[NF][V];
#pragma unroll
        for (int j = 0; j < NF; ++j) Vec<T>::unpack(r[h][j], e[j]);
#pragma unroll
        for (int q = 0; q < V; ++q) e[0][q] = fmaxf(e[0][q], 0.f);   // RIC: relu(x)
#pragma unroll
        for (int j = 0; j < NF; ++j)
#pragma unroll
          for (int q = 0; q < V; ++q) cs[j][q] += e[j][q];
#pragma unroll 2
        for (int i = 0; i < K; ++i) {
          float v[V];
          ldf<V>(&sBC[i * CB + lane * V], v);
          // coefficient of full cell 0 includes the gated skip (gate_i * emb_0, DynamicInteraction.py:57,66)
          const float c0 = sP[i * K] + sG[i];
#pragma unroll
          for (int q = 0; q < V; ++q) v[q] = fmaf(c0, e[0][q], v[q]);
#pragma unroll
          for (int j = 1; j < K; ++j) {
            if (is_bcast<K>(j)) continue;
            const float c = sP[i * K + j];
#pragma unroll
            for (int q = 0; q < V; ++q) v[q] = fmaf(c, e[full_idx<K>(j)][q], v[q]);
          }
          straw<T>(static_cast<T*>(const_cast<void*>(p.out.p[i])) + off, Vec<T>::pack(v));
        }
      }
    }
  }
  if (p.pooled) {
    // mean_L(out_i) = sum_j coef_ij * mean_L(e_j) + the broadcast cells' contribution (constant over L)
#pragma unroll
    for (int j = 0; j < NF; ++j) block_colsum<V>(cs[j], red, warp, lane);
    const float invL = 1.f / (float)p.L;
    if (active) {
      for (int i = warp; i < K; i += kWarps) {
        float v[V];
        ldf<V>(&sBC[i * CB + lane * V], v);
        const float c0 = (sP[i * K] + sG[i]) * invL;
#pragma unroll
        for (int q = 0; q < V; ++q) v[q] = fmaf(c0, cs[0][q], v[q]);
#pragma unroll
        for (int j = 1; j < K; ++j) {
          if (is_bcast<K>(j)) continue;
          const float c = sP[i * K + j] * invL;
#pragma unroll
          for (int q = 0; q < V; ++q) v[q] = fmaf(c, cs[full_idx<K>(j)][q], v[q]);
        }
        stf<V>(p.pooled + ((long long)i * p.B + b) * p.D + col, v);
      }
    }
  }
}

// ------------------------------------------------------------------ forward, final layer
template <typename T, int K>
__global__ void __launch_bounds__(kThreads, 4) agg_fwd_final_kernel(const AggP p) {
  constexpr int V = Vec<T>::V, CB = 32 * V, NF = num_full<K>();
  using Raw = typename Vec<T>::Raw;
  __shared__ float sP[K];
  __shared__ float sG[K];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = blockIdx.y;
  const long long col = (long long)blockIdx.x * CB + lane * V;
  if (threadIdx.x < K) {
    sP[threadIdx.x] = p.P[b * K + threadIdx.x];
    sG[threadIdx.x] = p.gate[b * K + threadIdx.x];
  }
  __syncthreads();
  if (col >= p.D) return;
  float S = 0.f;
#pragma unroll
  for (int j = 0; j < K; ++j) S += sP[j] + sG[j];
  const float inv = 1.f / S;
  float bc[V];
#pragma unroll
  for (int q = 0; q < V; ++q) bc[q] = 0.f;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    if (is_bcast<K>(j)) {
      float v[V];
      ldf<V>(static_cast<const float*>(p.bvec.p[j]) + b * p.D + col, v);
#pragma unroll
      for (int q = 0; q < V; ++q) bc[q] = fmaf(sP[j], v[q], bc[q]);
    }
  }
  const long long base = b * p.L * p.D + col;
  bool any_gate = false;
#pragma unroll
  for (int j = 1; j < K; ++j) any_gate = any_gate || sG[j] != 0.f;
  for (long long l = warp; l < p.L; l += 2 * kWarps) {
    const bool two = l + kWarps < p.L;
    Raw r[2][NF];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h == 1 && !two) break;
      const long long off = base + (l + h * kWarps) * p.D;
#pragma unroll
      for (int j = 0; j < K; ++j)
        if (!is_bcast<K>(j)) r[h][full_idx<K>(j)] = ldraw<T>(static_cast<const T*>(p.full.p[j]) + off);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h == 1 && !two) break;
      const long long off = base + (l + h * kWarps) * p.D;
      float e[NF][V], v[V];
#pragma unroll
      for (int j = 0; j < NF; ++j) Vec<T>::unpack(r[h][j], e[j]);
#pragma unroll
      for (int q = 0; q < V; ++q) {
        // cell 0: p_0 relu(x_0) + g_0 x_0  (full[0] is the raw layer input ref_wrd[0])
        float a = fmaf(sP[0], fmaxf(e[0][q], 0.f), fmaf(sG[0], e[0][q], bc[q]));
#pragma unroll
        for (int j = 1; j < K; ++j)
          if (!is_bcast<K>(j)) a = fmaf(sP[j], e[full_idx<K>(j)][q], a);
        v[q] = a;
      }
      if (any_gate) {
#pragma unroll
        for (int j = 1; j < K; ++j) {
          if (sG[j] != 0.f) {   // gated skip of cell j's input (p_j < 1e-4/K: rare)
            float x[V];
            Vec<T>::unpack(ldraw<T>(static_cast<const T*>(p.inputs.p[j]) + off), x);
#pragma unroll
            for (int q = 0; q < V; ++q) v[q] = fmaf(sG[j], x[q], v[q]);
          }
        }
      }
#pragma unroll
      for (int q = 0; q < V; ++q) v[q] *= inv;
      straw<T>(static_cast<T*>(const_cast<void*>(p.out.p[0])) + off, Vec<T>::pack(v));
    }
  }
}

// ------------------------------------------------------------------ backward, non-final
// g_i = d_out_i + d_pooled_i / L.   d_e_j = sum_i P_ij g_i (+ gate_i g_i for j = 0);
// dP_ij = sum_{l,d} g_i e_j;  d_bvec_j = sum_i P_ij sum_l g_i.
template <typename T, int K>
__global__ void __launch_bounds__(kThreads, 3) agg_bwd_kernel(const AggP p) {
  constexpr int V = Vec<T>::V, CB = 32 * V, NF = num_full<K>();
  using Raw = typename Vec<T>::Raw;
  __shared__ float sP[K * K];
  __shared__ float sG[K];
  __shared__ __align__(16) float sDpl[K * CB];
  // sum over this thread's rows of g_i, private slot per thread: [i][q / 4][thread] float4 (conflict-free).
  // It feeds both the broadcast cells' gradient and their dP entries; in registers it would cost K * V of them.
  __shared__ __align__(16) float4 sGs[K * (V / 4) * kThreads];
  __shared__ float sdP[kWarps][K * K];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = blockIdx.y;
  const long long col = (long long)blockIdx.x * CB + lane * V;
  const bool active = col < p.D;
  if (threadIdx.x < K * K) sP[threadIdx.x] = p.P[b * K * K + threadIdx.x];
  if (threadIdx.x < K) sG[threadIdx.x] = p.gate[b * K + threadIdx.x];
  const float invL = 1.f / (float)p.L;
  for (int idx = threadIdx.x; idx < K * CB; idx += kThreads) {
    const int i = idx / CB, c = idx % CB;
    const long long gc = (long long)blockIdx.x * CB + c;
    sDpl[idx] = (p.d_pooled && gc < p.D) ? p.d_pooled[((long long)i * p.B + b) * p.D + gc] * invL : 0.f;
  }
#pragma unroll
  for (int s = 0; s < K * (V / 4); ++s) sGs[s * kThreads + threadIdx.x] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  float dP[K][NF];     // dot products with the full cells
#pragma unroll
  for (int i = 0; i < K; ++i)
#pragma unroll
    for (int j = 0; j < NF; ++j) dP[i][j] = 0.f;
  if (active) {
    const long long base = b * p.L * p.D + col;
    for (long long l = warp; l < p.L; l += kWarps) {
      const long long off = base + l * p.D;
      Raw re[NF], rg[K];
#pragma unroll
      for (int j = 0; j < K; ++j)
        if (!is_bcast<K>(j)) re[full_idx<K>(j)] = ldraw<T>(static_cast<const T*>(p.full.p[j]) + off);
#pragma unroll
      for (int i = 0; i < K; ++i) rg[i] = ldraw<T>(static_cast<const T*>(p.d_out.p[i]) + off);
      float e[NF][V], de[NF][V];
#pragma unroll
      for (int j = 0; j < NF; ++j) {
        Vec<T>::unpack(re[j], e[j]);
#pragma unroll
        for (int q = 0; q < V; ++q) de[j][q] = 0.f;
      }
      float x0pos[V];
#pragma unroll
      for (int q = 0; q < V; ++q) {
        x0pos[q] = e[0][q] > 0.f ? 1.f : 0.f;
        e[0][q] = fmaxf(e[0][q], 0.f);
      }
#pragma unroll
      for (int i = 0; i < K; ++i) {
        float g[V];
        Vec<T>::unpack(rg[i], g);
#pragma unroll
        for (int q4 = 0; q4 < V / 4; ++q4) {
          const float4 d = *reinterpret_cast<const float4*>(&sDpl[i * CB + lane * V + q4 * 4]);
          g[q4 * 4] += d.x; g[q4 * 4 + 1] += d.y; g[q4 * 4 + 2] += d.z; g[q4 * 4 + 3] += d.w;
          float4 a = sGs[(i * (V / 4) + q4) * kThreads + threadIdx.x];
          a.x += g[q4 * 4]; a.y += g[q4 * 4 + 1]; a.z += g[q4 * 4 + 2]; a.w += g[q4 * 4 + 3];
          sGs[(i * (V / 4) + q4) * kThreads + threadIdx.x] = a;
        }
#pragma unroll
        for (int j = 0; j < K; ++j) {
          if (is_bcast<K>(j)) continue;
          const float w = sP[i * K + j] + (j == 0 ? sG[i] : 0.f);
          float dot = 0.f;
#pragma unroll
          for (int q = 0; q < V; ++q) {
            de[full_idx<K>(j)][q] = fmaf(w, g[q], de[full_idx<K>(j)][q]);
            dot = fmaf(g[q], e[full_idx<K>(j)][q], dot);
          }
          dP[i][full_idx<K>(j)] += dot;
        }
      }
#pragma unroll
      for (int q = 0; q < V; ++q) de[0][q] *= x0pos[q];
#pragma unroll
      for (int j = 0; j < K; ++j)
        if (!is_bcast<K>(j))
          straw<T>(static_cast<T*>(const_cast<void*>(p.d_full.p[j])) + off, Vec<T>::pack(de[full_idx<K>(j)]));
    }
  }
  __syncthreads();
  // broadcast cells: d_bvec_j[b,col] = sum_i P_ij * sum_l g_i;  dP_ij = sum_col bvec_j[col] * sum_l g_i[col].
  // Warp 0 sums the four warps' private slots of its lane's columns.
  float dPb[K][K - NF];
#pragma unroll
  for (int i = 0; i < K; ++i)
#pragma unroll
    for (int j = 0; j < K - NF; ++j) dPb[i][j] = 0.f;
  if (warp == 0 && active) {
    float eb[K - NF][V], db[K - NF][V];
#pragma unroll
    for (int j = 0; j < K; ++j) {
      if (!is_bcast<K>(j)) continue;
      ldf<V>(static_cast<const float*>(p.bvec.p[j]) + b * p.D + col, eb[bc_idx<K>(j)]);
#pragma unroll
      for (int q = 0; q < V; ++q) db[bc_idx<K>(j)][q] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < K; ++i) {
      float gs[V];
#pragma unroll
      for (int q4 = 0; q4 < V / 4; ++q4) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
          const float4 t = sGs[(i * (V / 4) + q4) * kThreads + w * 32 + lane];
          a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
        }
        gs[q4 * 4] = a.x; gs[q4 * 4 + 1] = a.y; gs[q4 * 4 + 2] = a.z; gs[q4 * 4 + 3] = a.w;
      }
#pragma unroll
      for (int j = 0; j < K; ++j) {
        if (!is_bcast<K>(j)) continue;
        float dot = 0.f;
#pragma unroll
        for (int q = 0; q < V; ++q) {
          db[bc_idx<K>(j)][q] = fmaf(sP[i * K + j], gs[q], db[bc_idx<K>(j)][q]);
          dot = fmaf(eb[bc_idx<K>(j)][q], gs[q], dot);
        }
        dPb[i][bc_idx<K>(j)] = dot;
      }
    }
#pragma unroll
    for (int j = 0; j < K; ++j)
      if (is_bcast<K>(j))
        stf<V>(static_cast<float*>(const_cast<void*>(p.d_bvec.p[j])) + b * p.D + col, db[bc_idx<K>(j)]);
  }
  // dP: warp reduce, then cross-warp via smem, one atomic per (i,j) per block
#pragma unroll
  for (int i = 0; i < K; ++i) {
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const float t = is_bcast<K>(j) ? dPb[i][bc_idx<K>(j)] : dP[i][full_idx<K>(j)];
      const float sm = warp_sum(t);
      if (lane == 0) sdP[warp][i * K + j] = sm;
    }
  }
  __syncthreads();
  if (threadIdx.x < K * K) {
    float sm = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) sm += sdP[w][threadIdx.x];
    atomicAdd(p.dP + b * K * K + threadIdx.x, sm);
  }
}

// ------------------------------------------------------------------ backward, final layer
// out = N / S,  N = sum_j (p_j e_j + g_j x_j),  S = sum_j (g_j + p_j)
// dN = d_out / S;  d_e_j = p_j dN;  dp_j = sum dN.e_j - sum dN.out.  The gated-skip gradient
// d_x_j = g_j dN (j >= 1) is produced by d2r_gate_skip_bwd only for samples whose gate is set.
template <typename T, int K>
__global__ void __launch_bounds__(kThreads, 3) agg_bwd_final_kernel(const AggP p) {
  constexpr int V = Vec<T>::V, CB = 32 * V, NF = num_full<K>();
  using Raw = typename Vec<T>::Raw;
  __shared__ float sP[K];
  __shared__ float sG[K];
  __shared__ __align__(16) float red[kWarps * CB];
  __shared__ float sdP[kWarps][K + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = blockIdx.y;
  const long long col = (long long)blockIdx.x * CB + lane * V;
  const bool active = col < p.D;
  if (threadIdx.x < K) {
    sP[threadIdx.x] = p.P[b * K + threadIdx.x];
    sG[threadIdx.x] = p.gate[b * K + threadIdx.x];
  }
  __syncthreads();
  float S = 0.f;
#pragma unroll
  for (int j = 0; j < K; ++j) S += sP[j] + sG[j];
  const float inv = 1.f / S;
  float eb[K - NF][V], dp[K + 1], dnsum[V], bcN[V];
#pragma unroll
  for (int q = 0; q < V; ++q) { dnsum[q] = 0.f; bcN[q] = 0.f; }
#pragma unroll
  for (int j = 0; j <= K; ++j) dp[j] = 0.f;   // dp[K] = sum dN . out
#pragma unroll
  for (int j = 0; j < K; ++j) {
    if (!is_bcast<K>(j)) continue;
#pragma unroll
    for (int q = 0; q < V; ++q) eb[bc_idx<K>(j)][q] = 0.f;
    if (active) {
      ldf<V>(static_cast<const float*>(p.bvec.p[j]) + b * p.D + col, eb[bc_idx<K>(j)]);
#pragma unroll
      for (int q = 0; q < V; ++q) bcN[q] = fmaf(sP[j], eb[bc_idx<K>(j)][q], bcN[q]);
    }
  }
  bool any_gate = false;
#pragma unroll
  for (int j = 1; j < K; ++j) any_gate = any_gate || sG[j] != 0.f;
  if (active) {
    const long long base = b * p.L * p.D + col;
    for (long long l = warp; l < p.L; l += 2 * kWarps) {
      const bool two = l + kWarps < p.L;
      Raw r[2][NF + 1];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (h == 1 && !two) break;
        const long long off = base + (l + h * kWarps) * p.D;
#pragma unroll
        for (int j = 0; j < K; ++j)
          if (!is_bcast<K>(j)) r[h][full_idx<K>(j)] = ldraw<T>(static_cast<const T*>(p.full.p[j]) + off);
        r[h][NF] = ldraw<T>(static_cast<const T*>(p.d_out.p[0]) + off);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (h == 1 && !two) break;
        const long long off = base + (l + h * kWarps) * p.D;
        float e[NF][V], dN[V], N[V], x0[V];
#pragma unroll
        for (int j = 0; j < NF; ++j) Vec<T>::unpack(r[h][j], e[j]);
        Vec<T>::unpack(r[h][NF], dN);
#pragma unroll
        for (int q = 0; q < V; ++q) {
          dN[q] *= inv;
          dnsum[q] += dN[q];
          x0[q] = e[0][q];
          e[0][q] = fmaxf(x0[q], 0.f);
          N[q] = fmaf(sG[0], x0[q], bcN[q]);
        }
#pragma unroll
        for (int j = 0; j < K; ++j) {
          float dot = 0.f;
          if (is_bcast<K>(j)) {
#pragma unroll
            for (int q = 0; q < V; ++q) dot = fmaf(dN[q], eb[bc_idx<K>(j)][q], dot);
          } else {
#pragma unroll
            for (int q = 0; q < V; ++q) {
              N[q] = fmaf(sP[j], e[full_idx<K>(j)][q], N[q]);
              dot = fmaf(dN[q], e[full_idx<K>(j)][q], dot);
            }
          }
          dp[j] += dot;
        }
        if (any_gate) {
#pragma unroll
          for (int j = 1; j < K; ++j) {
            if (sG[j] != 0.f) {
              float x[V];
              Vec<T>::unpack(ldraw<T>(static_cast<const T*>(p.inputs.p[j]) + off), x);
#pragma unroll
              for (int q = 0; q < V; ++q) N[q] = fmaf(sG[j], x[q], N[q]);
            }
          }
        }
        {
          float dot = 0.f;
#pragma unroll
          for (int q = 0; q < V; ++q) dot = fmaf(dN[q], N[q] * inv, dot);
          dp[K] += dot;
        }
#pragma unroll
        for (int j = 0; j < K; ++j) {
          if (is_bcast<K>(j)) continue;
          float v[V];
#pragma unroll
          for (int q = 0; q < V; ++q)
            v[q] = j == 0 ? dN[q] * (sP[0] * (x0[q] > 0.f ? 1.f : 0.f) + sG[0]) : sP[j] * dN[q];
          straw<T>(static_cast<T*>(const_cast<void*>(p.d_full.p[j])) + off, Vec<T>::pack(v));
        }
      }
    }
  }
  block_colsum<V>(dnsum, red, warp, lane);
  if (warp == 0 && active) {
#pragma unroll
    for (int j = 0; j < K; ++j) {
      if (is_bcast<K>(j)) {
        float v[V];
#pragma unroll
        for (int q = 0; q < V; ++q) v[q] = sP[j] * dnsum[q];
        stf<V>(static_cast<float*>(const_cast<void*>(p.d_bvec.p[j])) + b * p.D + col, v);
      }
    }
  }
#pragma unroll
  for (int j = 0; j <= K; ++j) {
    const float sm = warp_sum(dp[j]);
    if (lane == 0) sdP[warp][j] = sm;
  }
  __syncthreads();
  if (threadIdx.x < K) {
    float sm = 0.f, t = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      sm += sdP[w][threadIdx.x];
      t += sdP[w][K];
    }
    atomicAdd(p.dP + b * K + threadIdx.x, sm - t);
  }
}

// grid (B, K-1, chunks): dx_j[b] (+)= gate[b,j] / S[b] * d_out[b] for the samples whose gate is set; blocks of
// un-gated samples exit immediately (the common case), or zero-fill when `zero_fill` (no prior gradient).
template <typename T>
__global__ void __launch_bounds__(256) gate_skip_bwd_kernel(const T* __restrict__ d_out, const float* __restrict__ P,
                                                            const float* __restrict__ gate, d2r_ptr8 dx, int K,
                                                            long long LD, int accumulate_mask) {
  const long long b = blockIdx.x;
  const int j = blockIdx.y + 1;
  T* dst = static_cast<T*>(const_cast<void*>(dx.p[j]));
  if (dst == nullptr) return;
  const float g = gate[b * K + j];
  const bool acc = (accumulate_mask >> j) & 1;
  if (g == 0.f && acc) return;
  float S = 0.f;
  for (int c = 0; c < K; ++c) S += P[b * K + c] + gate[b * K + c];
  const float w = g / S;
  const long long nvec = LD / 8;
  for (long long i = blockIdx.z * 256 + threadIdx.x; i < nvec; i += (long long)gridDim.z * 256) {
    float v[8], o[8];
    load8(d_out + b * LD + i * 8, v);
    if (acc) {
      load8(dst + b * LD + i * 8, o);
#pragma unroll
      for (int q = 0; q < 8; ++q) o[q] = fmaf(w, v[q], o[q]);
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) o[q] = w * v[q];
    }
    store8(dst + b * LD + i * 8, o);
  }
}

int fill_common(AggP& q, const d2r_agg_args& a) {
  D2R_CHECK_ARG(a.K == 4 || a.K == 6, "aggregate: K must be 4 or 6 (got %d)", a.K);
  D2R_CHECK_ARG(a.final_layer ? a.n_out == 1 : a.n_out == a.K, "aggregate: n_out must be K (or 1 in the final layer)");
  D2R_CHECK_ARG(a.B > 0 && a.L > 0 && a.D > 0 && a.B <= 65535, "aggregate: bad shape");
  D2R_CHECK_ARG(a.D % (a.dtype == D2R_BF16 ? 8 : 4) == 0, "aggregate: D must be a multiple of 8 (bf16) / 4 (fp32): 16-byte accesses");
  for (int j = 0; j < a.K; ++j) {
    const bool bc = a.K == 6 ? is_bcast<6>(j) : is_bcast<4>(j);
    D2R_CHECK_ARG(bc ? (a.bvec.p[j] != nullptr && a.full.p[j] == nullptr)
                     : (a.full.p[j] != nullptr && a.bvec.p[j] == nullptr),
                  "aggregate: cell %d must be a %s cell (emb_lst order ric,glac,imrc,cmrc,crcmc,gesc)", j,
                  bc ? "broadcast [B,D]" : "full [B,L,D]");
  }
  q.K = a.K; q.n_out = a.n_out; q.B = a.B; q.L = a.L; q.D = a.D;
  q.full = a.full; q.bvec = a.bvec; q.inputs = a.inputs; q.out = a.out;
  q.P = a.P; q.gate = a.gate; q.pooled = a.pooled;
  q.d_pooled = nullptr; q.dP = nullptr;
  return D2R_OK;
}

}  // namespace

extern "C" {

int d2r_aggregate_fwd(const d2r_agg_args* a, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  D2R_CHECK_ARG(a != nullptr, "aggregate: null args");
  AggP q{};
  if (int rc = fill_common(q, *a)) return rc;
  const int cb = a->dtype == D2R_BF16 ? 256 : 128;     // columns per block: 32 lanes x 16 bytes
  dim3 grid((unsigned)((a->D + cb - 1) / cb), (unsigned)a->B);
#define D2R_AGG_LAUNCH(KERN)                                                         \
  do {                                                                               \
    if (a->K == 6) { D2R_DISPATCH_DTYPE(a->dtype, T, KERN<T, 6><<<grid, kThreads, 0, st>>>(q)); } \
    else           { D2R_DISPATCH_DTYPE(a->dtype, T, KERN<T, 4><<<grid, kThreads, 0, st>>>(q)); } \
  } while (0)
  if (a->final_layer) D2R_AGG_LAUNCH(agg_fwd_final_kernel);
  else D2R_AGG_LAUNCH(agg_fwd_kernel);
  count_launch();
  return check_launch("agg_fwd_kernel");
}

int d2r_aggregate_bwd(const d2r_agg_bwd_args* a, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  D2R_CHECK_ARG(a != nullptr && a->dP != nullptr, "aggregate_bwd: null args");
  AggP q{};
  if (int rc = fill_common(q, a->fwd)) return rc;
  q.d_out = a->d_out; q.d_full = a->d_full; q.d_bvec = a->d_bvec;
  q.d_pooled = a->d_pooled; q.dP = a->dP;
  const d2r_agg_args* f = &a->fwd;
  for (int j = 0; j < f->K; ++j) {
    const bool bc = f->K == 6 ? is_bcast<6>(j) : is_bcast<4>(j);
    D2R_CHECK_ARG(bc ? a->d_bvec.p[j] != nullptr : a->d_full.p[j] != nullptr, "aggregate_bwd: missing gradient buffer %d", j);
  }
  D2R_CUDA_OK(cudaMemsetAsync(a->dP, 0, sizeof(float) * (size_t)f->B * f->n_out * f->K, st));
  const int cb = f->dtype == D2R_BF16 ? 256 : 128;
  dim3 grid((unsigned)((f->D + cb - 1) / cb), (unsigned)f->B);
#define D2R_AGGB_LAUNCH(KERN)                                                         \
  do {                                                                                \
    if (f->K == 6) { D2R_DISPATCH_DTYPE(f->dtype, T, KERN<T, 6><<<grid, kThreads, 0, st>>>(q)); } \
    else           { D2R_DISPATCH_DTYPE(f->dtype, T, KERN<T, 4><<<grid, kThreads, 0, st>>>(q)); } \
  } while (0)
  if (f->final_layer) D2R_AGGB_LAUNCH(agg_bwd_final_kernel);
  else D2R_AGGB_LAUNCH(agg_bwd_kernel);
  count_launch();
  return check_launch("agg_bwd_kernel");
}

int d2r_gate_skip_bwd(const void* d_out, const float* P, const float* gate, d2r_ptr8 dx, int32_t K, int64_t B,
                      int64_t L, int64_t D, int32_t dtype, int32_t accumulate_mask, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  D2R_CHECK_ARG(K >= 2 && K <= 8 && B > 0 && (L * D) % 8 == 0 && B <= 0x7fffffff, "gate_skip_bwd: bad shape");
  dim3 grid((unsigned)B, (unsigned)(K - 1), 8);
  D2R_DISPATCH_DTYPE(dtype, T,
                     gate_skip_bwd_kernel<T><<<grid, 256, 0, st>>>((const T*)d_out, P, gate, dx, K, L * D, accumulate_mask));
  count_launch();
  return check_launch("gate_skip_bwd_kernel");
}

}  // extern "C"
}  // namespace d2r
