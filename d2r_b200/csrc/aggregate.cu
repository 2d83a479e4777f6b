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
// Grid (D/128, B): a block owns 128 columns of one sample and walks all L rows (4 warps, two rows in
// flight per warp), so the reductions over L (pooled means, broadcast-cell gradients, dP) need no
// second pass.  The set of broadcast cells is a compile-time mask (cells 1 and 5 of emb_lst), which
// keeps the per-thread state in registers.
#include "common.cuh"
#include "ptx.cuh"

namespace d2r {
namespace {

constexpr int kWarps = 4;
constexpr int kThreads = kWarps * 32;
constexpr int kCols = 128;   // columns per block (lane owns 4 consecutive)

__device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
__device__ __forceinline__ void load4(const __nv_bfloat16* p, float (&v)[4]) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
  const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void store4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(__nv_bfloat16* p, const float (&v)[4]) {
  uint2 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
  h[0] = __floats2bfloat162_rn(v[0], v[1]);
  h[1] = __floats2bfloat162_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = u;
}

// broadcast cells of emb_lst: GLAC (1) and, with six cells, GESC (5)
template <int K> __host__ __device__ constexpr bool is_bcast(int j) { return j == 1 || (K == 6 && j == 5); }

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

// cross-warp sum of per-thread partials v[4] for this block's 128 columns; red: [kWarps][kCols]
__device__ __forceinline__ void block_colsum4(float (&v)[4], float (*red)[kCols], int warp, int lane) {
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 4; ++q) red[warp][lane * 4 + q] = v[q];
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += red[w][lane * 4 + q];
    v[q] = s;
  }
}

// ------------------------------------------------------------------ forward, non-final
template <typename T, int K>
__global__ void __launch_bounds__(kThreads) agg_fwd_kernel(const AggP p) {
  __shared__ float sP[K * K];
  __shared__ float sG[K];
  __shared__ float red[kWarps][kCols];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = blockIdx.y;
  const long long col = (long long)blockIdx.x * kCols + lane * 4;
  if (threadIdx.x < K * K) sP[threadIdx.x] = p.P[b * K * K + threadIdx.x];
  if (threadIdx.x < K) sG[threadIdx.x] = p.gate[b * K + threadIdx.x];
  __syncthreads();
  const bool active = col < p.D;
  // contribution of the broadcast cells to output i is constant over the rows
  float bc[K][4], pool[K][4];
#pragma unroll
  for (int i = 0; i < K; ++i)
#pragma unroll
    for (int q = 0; q < 4; ++q) { bc[i][q] = 0.f; pool[i][q] = 0.f; }
  if (active) {
#pragma unroll
    for (int j = 0; j < K; ++j) {
      if (is_bcast<K>(j)) {
        float v[4];
        load4(static_cast<const float*>(p.bvec.p[j]) + b * p.D + col, v);
#pragma unroll
        for (int i = 0; i < K; ++i)
#pragma unroll
          for (int q = 0; q < 4; ++q) bc[i][q] = fmaf(sP[i * K + j], v[q], bc[i][q]);
      }
    }
    const long long base = b * p.L * p.D + col;
#pragma unroll 2
    for (long long l = warp; l < p.L; l += kWarps) {
      const long long off = base + l * p.D;
      float e[K][4];
#pragma unroll
      for (int j = 0; j < K; ++j)
        if (!is_bcast<K>(j)) load4(static_cast<const T*>(p.full.p[j]) + off, e[j]);
#pragma unroll
      for (int q = 0; q < 4; ++q) e[0][q] = fmaxf(e[0][q], 0.f);   // RIC: relu(x)
#pragma unroll
      for (int i = 0; i < K; ++i) {
        float v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float a = fmaf(sG[i], e[0][q], bc[i][q]);
#pragma unroll
          for (int j = 0; j < K; ++j)
            if (!is_bcast<K>(j)) a = fmaf(sP[i * K + j], e[j][q], a);
          v[q] = a;
          pool[i][q] += a;
        }
        store4(static_cast<T*>(const_cast<void*>(p.out.p[i])) + off, v);
      }
    }
  }
  if (p.pooled) {
    const float invL = 1.f / (float)p.L;
#pragma unroll
    for (int i = 0; i < K; ++i) {
      block_colsum4(pool[i], red, warp, lane);
      if (warp == 0 && active) {
        float v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = pool[i][q] * invL;
        store4(p.pooled + ((long long)i * p.B + b) * p.D + col, v);
      }
    }
  }
}

// ------------------------------------------------------------------ forward, final layer
template <typename T, int K>
__global__ void __launch_bounds__(kThreads) agg_fwd_final_kernel(const AggP p) {
  __shared__ float sP[K];
  __shared__ float sG[K];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = blockIdx.y;
  const long long col = (long long)blockIdx.x * kCols + lane * 4;
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
  float bc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < K; ++j) {
    if (is_bcast<K>(j)) {
      float v[4];
      load4(static_cast<const float*>(p.bvec.p[j]) + b * p.D + col, v);
#pragma unroll
      for (int q = 0; q < 4; ++q) bc[q] = fmaf(sP[j], v[q], bc[q]);
    }
  }
  const long long base = b * p.L * p.D + col;
#pragma unroll 2
  for (long long l = warp; l < p.L; l += kWarps) {
    const long long off = base + l * p.D;
    float e[K][4];
#pragma unroll
    for (int j = 0; j < K; ++j)
      if (!is_bcast<K>(j)) load4(static_cast<const T*>(p.full.p[j]) + off, e[j]);
    float v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      // cell 0: p_0 relu(x_0) + g_0 x_0  (full[0] is the raw layer input ref_wrd[0])
      float a = fmaf(sP[0], fmaxf(e[0][q], 0.f), fmaf(sG[0], e[0][q], bc[q]));
#pragma unroll
      for (int j = 1; j < K; ++j)
        if (!is_bcast<K>(j)) a = fmaf(sP[j], e[j][q], a);
      v[q] = a;
    }
#pragma unroll
    for (int j = 1; j < K; ++j) {
      if (sG[j] != 0.f) {   // gated skip of cell j's input (p_j < 1e-4/K: rare)
        float x[4];
        load4(static_cast<const T*>(p.inputs.p[j]) + off, x);
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = fmaf(sG[j], x[q], v[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] *= inv;
    store4(static_cast<T*>(const_cast<void*>(p.out.p[0])) + off, v);
  }
}

// ------------------------------------------------------------------ backward, non-final
// g_i = d_out_i + d_pooled_i / L.   d_e_j = sum_i P_ij g_i (+ gate_i g_i for j = 0);
// dP_ij = sum_{l,d} g_i e_j;  d_bvec_j = sum_i P_ij sum_l g_i.
template <typename T, int K>
__global__ void __launch_bounds__(kThreads) agg_bwd_kernel(const AggP p) {
  __shared__ float sP[K * K];
  __shared__ float sG[K];
  __shared__ float sDpl[K][kCols];
  __shared__ float red[kWarps][kCols];
  __shared__ float sdP[kWarps][K * K];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = blockIdx.y;
  const long long col = (long long)blockIdx.x * kCols + lane * 4;
  const bool active = col < p.D;
  if (threadIdx.x < K * K) sP[threadIdx.x] = p.P[b * K * K + threadIdx.x];
  if (threadIdx.x < K) sG[threadIdx.x] = p.gate[b * K + threadIdx.x];
  const float invL = 1.f / (float)p.L;
  for (int idx = threadIdx.x; idx < K * kCols; idx += kThreads) {
    const int i = idx / kCols, c = idx % kCols;
    const long long gc = (long long)blockIdx.x * kCols + c;
    sDpl[i][c] = (p.d_pooled && gc < p.D) ? p.d_pooled[((long long)i * p.B + b) * p.D + gc] * invL : 0.f;
  }
  __syncthreads();
  float eb[K][4];      // broadcast cells' values (constant over the rows)
  float gsum[K][4];    // sum over rows of g_i
  float dP[K][K];
#pragma unroll
  for (int j = 0; j < K; ++j) {
#pragma unroll
    for (int q = 0; q < 4; ++q) { eb[j][q] = 0.f; gsum[j][q] = 0.f; }
#pragma unroll
    for (int i = 0; i < K; ++i) dP[i][j] = 0.f;
    if (active && is_bcast<K>(j)) load4(static_cast<const float*>(p.bvec.p[j]) + b * p.D + col, eb[j]);
  }
  if (active) {
    const long long base = b * p.L * p.D + col;
    for (long long l = warp; l < p.L; l += kWarps) {
      const long long off = base + l * p.D;
      float e[K][4], de[K][4];
#pragma unroll
      for (int j = 0; j < K; ++j) {
        if (!is_bcast<K>(j)) load4(static_cast<const T*>(p.full.p[j]) + off, e[j]);
#pragma unroll
        for (int q = 0; q < 4; ++q) de[j][q] = 0.f;
      }
      float x0pos[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        x0pos[q] = e[0][q] > 0.f ? 1.f : 0.f;
        e[0][q] = fmaxf(e[0][q], 0.f);
      }
#pragma unroll
      for (int i = 0; i < K; ++i) {
        float g[4];
        load4(static_cast<const T*>(p.d_out.p[i]) + off, g);
#pragma unroll
        for (int q = 0; q < 4; ++q) g[q] += sDpl[i][lane * 4 + q];
#pragma unroll
        for (int j = 0; j < K; ++j) {
          float dot = 0.f;
          if (is_bcast<K>(j)) {
#pragma unroll
            for (int q = 0; q < 4; ++q) dot = fmaf(g[q], eb[j][q], dot);
          } else {
            const float w = sP[i * K + j] + (j == 0 ? sG[i] : 0.f);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              de[j][q] = fmaf(w, g[q], de[j][q]);
              dot = fmaf(g[q], e[j][q], dot);
            }
          }
          dP[i][j] += dot;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) gsum[i][q] += g[q];
      }
#pragma unroll
      for (int j = 0; j < K; ++j) {
        if (!is_bcast<K>(j)) {
          if (j == 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q) de[0][q] *= x0pos[q];
          }
          store4(static_cast<T*>(const_cast<void*>(p.d_full.p[j])) + off, de[j]);
        }
      }
    }
  }
  // broadcast-cell gradients: d_bvec_j[b,col] = sum_i P_ij * sum_l g_i
#pragma unroll
  for (int i = 0; i < K; ++i) block_colsum4(gsum[i], red, warp, lane);
  if (warp == 0 && active) {
#pragma unroll
    for (int j = 0; j < K; ++j) {
      if (is_bcast<K>(j)) {
        float v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float a = 0.f;
#pragma unroll
          for (int i = 0; i < K; ++i) a = fmaf(sP[i * K + j], gsum[i][q], a);
          v[q] = a;
        }
        store4(static_cast<float*>(const_cast<void*>(p.d_bvec.p[j])) + b * p.D + col, v);
      }
    }
  }
  // dP: warp reduce, then cross-warp via smem, one atomic per (i,j) per block
#pragma unroll
  for (int i = 0; i < K; ++i)
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const float s = warp_sum(dP[i][j]);
      if (lane == 0) sdP[warp][i * K + j] = s;
    }
  __syncthreads();
  if (threadIdx.x < K * K) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += sdP[w][threadIdx.x];
    atomicAdd(p.dP + b * K * K + threadIdx.x, s);
  }
}

// ------------------------------------------------------------------ backward, final layer
// out = N / S,  N = sum_j (p_j e_j + g_j x_j),  S = sum_j (g_j + p_j)
// dN = d_out / S;  d_e_j = p_j dN;  dp_j = sum dN.e_j - sum dN.out.  The gated-skip gradient
// d_x_j = g_j dN (j >= 1) is produced by d2r_gate_skip_bwd only for samples whose gate is set.
template <typename T, int K>
__global__ void __launch_bounds__(kThreads) agg_bwd_final_kernel(const AggP p) {
  __shared__ float sP[K];
  __shared__ float sG[K];
  __shared__ float red[kWarps][kCols];
  __shared__ float sdP[kWarps][K + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = blockIdx.y;
  const long long col = (long long)blockIdx.x * kCols + lane * 4;
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
  float eb[K][4], dp[K + 1], dnsum[4], bcN[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) { dnsum[q] = 0.f; bcN[q] = 0.f; }
#pragma unroll
  for (int j = 0; j < K; ++j) {
    dp[j] = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) eb[j][q] = 0.f;
    if (active && is_bcast<K>(j)) {
      load4(static_cast<const float*>(p.bvec.p[j]) + b * p.D + col, eb[j]);
#pragma unroll
      for (int q = 0; q < 4; ++q) bcN[q] = fmaf(sP[j], eb[j][q], bcN[q]);
    }
  }
  dp[K] = 0.f;   // sum dN . out
  if (active) {
    const long long base = b * p.L * p.D + col;
#pragma unroll 2
    for (long long l = warp; l < p.L; l += kWarps) {
      const long long off = base + l * p.D;
      float e[K][4];
#pragma unroll
      for (int j = 0; j < K; ++j)
        if (!is_bcast<K>(j)) load4(static_cast<const T*>(p.full.p[j]) + off, e[j]);
      float dN[4], N[4], x0[4];
      load4(static_cast<const T*>(p.d_out.p[0]) + off, dN);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
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
          for (int q = 0; q < 4; ++q) dot = fmaf(dN[q], eb[j][q], dot);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            N[q] = fmaf(sP[j], e[j][q], N[q]);
            dot = fmaf(dN[q], e[j][q], dot);
          }
        }
        dp[j] += dot;
      }
#pragma unroll
      for (int j = 1; j < K; ++j) {
        if (sG[j] != 0.f) {
          float x[4];
          load4(static_cast<const T*>(p.inputs.p[j]) + off, x);
#pragma unroll
          for (int q = 0; q < 4; ++q) N[q] = fmaf(sG[j], x[q], N[q]);
        }
      }
      {
        float dot = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) dot = fmaf(dN[q], N[q] * inv, dot);
        dp[K] += dot;
      }
#pragma unroll
      for (int j = 0; j < K; ++j) {
        if (!is_bcast<K>(j)) {
          float v[4];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            v[q] = j == 0 ? dN[q] * (sP[0] * (x0[q] > 0.f ? 1.f : 0.f) + sG[0]) : sP[j] * dN[q];
          store4(static_cast<T*>(const_cast<void*>(p.d_full.p[j])) + off, v);
        }
      }
    }
  }
  block_colsum4(dnsum, red, warp, lane);
  if (warp == 0 && active) {
#pragma unroll
    for (int j = 0; j < K; ++j) {
      if (is_bcast<K>(j)) {
        float v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = sP[j] * dnsum[q];
        store4(static_cast<float*>(const_cast<void*>(p.d_bvec.p[j])) + b * p.D + col, v);
      }
    }
  }
#pragma unroll
  for (int j = 0; j <= K; ++j) {
    const float s = warp_sum(dp[j]);
    if (lane == 0) sdP[warp][j] = s;
  }
  __syncthreads();
  if (threadIdx.x < K) {
    float s = 0.f, t = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      s += sdP[w][threadIdx.x];
      t += sdP[w][K];
    }
    atomicAdd(p.dP + b * K + threadIdx.x, s - t);
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
  D2R_CHECK_ARG(a.B > 0 && a.L > 0 && a.D > 0 && a.D % 4 == 0 && a.B <= 65535, "aggregate: bad shape");
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
  dim3 grid((unsigned)((a->D + kCols - 1) / kCols), (unsigned)a->B);
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
  dim3 grid((unsigned)((f->D + kCols - 1) / kCols), (unsigned)f->B);
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
