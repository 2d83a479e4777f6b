// HBM-bound elementwise / row-reduction kernels of the cells: dtype casts, activation
// backward + bias gradient, row softmax, row L2-norm, FiLM modulation, squared-difference
// backward, GESC gate.  All use 16-byte vector accesses on the contiguous dimension.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace d2r {

namespace {

constexpr int kThreads = 256;

inline unsigned blocks_for(long long work_items, int per_block) {
  long long b = (work_items + per_block - 1) / per_block;
  const long long cap = 148LL * 32;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ------------------------------------------------------------------ cast / axpby
template <typename TI, typename TO>
__global__ void cast_kernel(const TI* __restrict__ x, TO* __restrict__ y, long long n) {
  const long long nvec = n / 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    float v[8];
    load8(x + i * 8, v);
    store8(y + i * 8, v);
  }
  for (long long i = nvec * 8 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    Elem<TO>::st(y + i, Elem<TI>::ld(x + i));
}

template <typename T>
__global__ void axpby_kernel(const T* __restrict__ x, const T* __restrict__ z, float a, float b, T* __restrict__ y,
                             long long n) {
  const long long nvec = n / 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    float vx[8], vz[8];
    load8(x + i * 8, vx);
    if (z) load8(z + i * 8, vz);
#pragma unroll
    for (int j = 0; j < 8; ++j) vx[j] = a * vx[j] + (z ? b * vz[j] : 0.f);
    store8(y + i * 8, vx);
  }
  for (long long i = nvec * 8 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    Elem<T>::st(y + i, a * Elem<T>::ld(x + i) + (z ? b * Elem<T>::ld(z + i) : 0.f));
}

// ------------------------------------------------------------------ act backward + bias gradient
// block = 8 warps; a block owns 256 columns x 64 rows; lane owns 8 consecutive columns; every warp keeps
// 8 independent 16/32-byte row loads in flight (raw vectors, unpacked one row at a time).
constexpr int kRowsPerBlock = 64;
template <typename T> struct Raw8;
template <> struct Raw8<__nv_bfloat16> {
  uint4 u;
  __device__ __forceinline__ void ld(const __nv_bfloat16* p) { u = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void get(float (&v)[8]) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x;
      v[2 * i + 1] = f.y;
    }
  }
};
template <> struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void ld(const float* p) {
    a = *reinterpret_cast<const float4*>(p);
    b = *reinterpret_cast<const float4*>(p + 4);
  }
  __device__ __forceinline__ void get(float (&v)[8]) const {
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};

template <typename T, bool ACT>
__global__ void __launch_bounds__(kThreads) bias_act_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y,
                                                                int act, T* __restrict__ dz, float* __restrict__ db,
                                                                long long rows, int cols, long long ld,
                                                                int rows_per_block) {
  __shared__ float red[8][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = blockIdx.x * 256 + lane * 8;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (col < cols) {
    constexpr int U = ACT ? 4 : 8;
    for (long long rb = r0 + warp * U; rb < r1; rb += 8 * U) {
      Raw8<T> g[U], yv[ACT ? U : 1];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long r = rb + u;
        if (r < r1) {
          g[u].ld(dy + r * ld + col);
          if constexpr (ACT) yv[u].ld(y + r * ld + col);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long r = rb + u;
        if (r < r1) {
          float gv[8];
          g[u].get(gv);
          if constexpr (ACT) {
            float yf[8];
            yv[u].get(yf);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              gv[j] = act == D2R_ACT_RELU ? (yf[j] > 0.f ? gv[j] : 0.f) : gv[j] * (1.f - yf[j] * yf[j]);
            if (dz) store8(dz + r * ld + col, gv);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += gv[j];
        }
      }
    }
  }
  if (db) {
#pragma unroll
    for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = acc[j];
    __syncthreads();
    const int c = threadIdx.x;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][c];
    if (blockIdx.x * 256 + c < cols) atomicAdd(db + blockIdx.x * 256 + c, s);
  }
}

// ------------------------------------------------------------------ row softmax (one warp per row)
template <typename TX, typename TY, int PER>
__global__ void __launch_bounds__(kThreads) softmax_fwd_kernel(const TX* __restrict__ x, long long ldx,
                                                               TY* __restrict__ y, long long ldy, long long rows,
                                                               int cols, float scale) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const TX* xr = x + row * ldx;
  float v[PER];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = lane + 32 * i;
    v[i] = c < cols ? scale * Elem<TX>::ld(xr + c) : -INFINITY;
    mx = fmaxf(mx, v[i]);
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    v[i] = (lane + 32 * i) < cols ? __expf(v[i] - mx) : 0.f;
    sum += v[i];
  }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  TY* yr = y + row * ldy;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = lane + 32 * i;
    if (c < cols) Elem<TY>::st(yr + c, v[i] * inv);
  }
}

template <typename TY, typename TD, typename TX, int PER>
__global__ void __launch_bounds__(kThreads) softmax_bwd_kernel(const TY* __restrict__ y, long long ldy,
                                                               const TD* __restrict__ dy, long long lddy,
                                                               TX* __restrict__ dx, long long lddx, long long rows,
                                                               int cols, float scale) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float yv[PER], gv[PER];
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = lane + 32 * i;
    yv[i] = c < cols ? Elem<TY>::ld(y + row * ldy + c) : 0.f;
    gv[i] = c < cols ? Elem<TD>::ld(dy + row * lddy + c) : 0.f;
    dot += yv[i] * gv[i];
  }
  dot = warp_sum(dot);
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = lane + 32 * i;
    if (c < cols) Elem<TX>::st(dx + row * lddx + c, scale * yv[i] * (gv[i] - dot));
  }
}

// ------------------------------------------------------------------ row L2 norm (one warp per row)
template <typename T>
__global__ void __launch_bounds__(kThreads) l2norm_fwd_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                              float* __restrict__ rnorm, long long rows, int cols) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const T* xr = x + row * cols;
  float ss = 0.f;
  for (int c = lane * 8; c < cols; c += 256) {
    float v[8];
    load8(xr + c, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) ss += v[j] * v[j];
  }
  ss = warp_sum(ss);
  const float r = 1.f / (sqrtf(ss) + 1e-8f);
  if (lane == 0) rnorm[row] = r;
  for (int c = lane * 8; c < cols; c += 256) {
    float v[8];
    load8(xr + c, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= r;
    store8(y + row * cols + c, v);
  }
}

// y = x r, r = 1/(n+eps):  dx = r dy - y (dy.y)/n,  n = 1/r - eps
template <typename T>
__global__ void __launch_bounds__(kThreads) l2norm_bwd_kernel(const T* __restrict__ y, const T* __restrict__ dy,
                                                              const float* __restrict__ rnorm, T* __restrict__ dx,
                                                              long long rows, int cols) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float dot = 0.f;
  for (int c = lane * 8; c < cols; c += 256) {
    float a[8], b[8];
    load8(y + row * cols + c, a);
    load8(dy + row * cols + c, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) dot += a[j] * b[j];
  }
  dot = warp_sum(dot);
  const float r = rnorm[row];
  const float n = fmaxf(1.f / r - 1e-8f, 1e-30f);
  const float k = dot / n;
  for (int c = lane * 8; c < cols; c += 256) {
    float a[8], b[8];
    load8(y + row * cols + c, a);
    load8(dy + row * cols + c, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] = r * b[j] - a[j] * k;
    store8(dx + row * cols + c, b);
  }
}

// ------------------------------------------------------------------ FiLM
template <typename T>
__global__ void film_fwd_kernel(const T* __restrict__ x, const T* __restrict__ st, T* __restrict__ m,
                                long long rows, int cols) {
  const int vpr = cols / 8;
  const long long total = rows * vpr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vpr;
    const int c = (int)(i % vpr) * 8;
    float xv[8], s[8], t[8];
    load8(x + r * cols + c, xv);
    load8(st + r * 2 * cols + c, s);
    load8(st + r * 2 * cols + cols + c, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) xv[j] = xv[j] * s[j] + t[j];
    store8(m + r * cols + c, xv);
  }
}

template <typename T>
__global__ void film_bwd_kernel(const T* __restrict__ dm, const T* __restrict__ x, const T* __restrict__ st,
                                const T* __restrict__ add, T* __restrict__ dx, T* __restrict__ dst, long long rows, int cols) {
  const int vpr = cols / 8;
  const long long total = rows * vpr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vpr;
    const int c = (int)(i % vpr) * 8;
    float g[8], xv[8], s[8], o[8];
    load8(dm + r * cols + c, g);
    load8(x + r * cols + c, xv);
    load8(st + r * 2 * cols + c, s);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = g[j] * s[j];
    if (add) {
      float ad[8];
      load8(add + r * cols + c, ad);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += ad[j];
    }
    store8(dx + r * cols + c, o);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = g[j] * xv[j] * (1.f - s[j] * s[j]);   // through tanh
    store8(dst + r * 2 * cols + c, o);
    store8(dst + r * 2 * cols + cols + c, g);
  }
}

// g = 2 d dsq  (dx = g, dc = -g)
template <typename T>
__global__ void sqdiff_bwd_kernel(const T* __restrict__ dsq, const T* __restrict__ d, const T* __restrict__ add,
                                  T* __restrict__ g, T* __restrict__ gx, long long n) {
  const long long nvec = n / 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    float a[8], b[8];
    load8(dsq + i * 8, a);
    load8(d + i * 8, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = 2.f * a[j] * b[j];
    store8(g + i * 8, a);
    if (gx) {
      if (add) {
        load8(add + i * 8, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] += b[j];
      }
      store8(gx + i * 8, a);
    }
  }
  for (long long i = nvec * 8 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
  {
    const float v = 2.f * Elem<T>::ld(dsq + i) * Elem<T>::ld(d + i);
    Elem<T>::st(g + i, v);
    if (gx) Elem<T>::st(gx + i, v + (add ? Elem<T>::ld(add + i) : 0.f));
  }
}

template <typename T>
__global__ void mul_kernel(const T* __restrict__ x, const T* __restrict__ z, float alpha, T* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    Elem<T>::st(y + i, alpha * Elem<T>::ld(x + i) * Elem<T>::ld(z + i));
}

// ------------------------------------------------------------------ GESC gate: block per row
__device__ __forceinline__ float block_reduce(float v, float* sm, bool is_max) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  float r = sm[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = is_max ? fmaxf(r, sm[w]) : r + sm[w];
  return r;
}

__global__ void __launch_bounds__(kThreads) gate_fuse_fwd_kernel(const float* __restrict__ gl,
                                                                 const float* __restrict__ t,
                                                                 const float* __restrict__ im, float* __restrict__ g,
                                                                 float* __restrict__ out, int D) {
  __shared__ float sm[8];
  const long long b = blockIdx.x;
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < D; c += kThreads) mx = fmaxf(mx, gl[b * D + c]);
  mx = block_reduce(mx, sm, true);
  float sum = 0.f;
  for (int c = threadIdx.x; c < D; c += kThreads) sum += __expf(gl[b * D + c] - mx);
  sum = block_reduce(sum, sm, false);
  const float inv = 1.f / sum;
  for (int c = threadIdx.x; c < D; c += kThreads) {
    const float gv = __expf(gl[b * D + c] - mx) * inv;
    g[b * D + c] = gv;
    out[b * D + c] = gv * t[b * D + c] + (1.f - gv) * im[b * D + c];
  }
}

__global__ void __launch_bounds__(kThreads) gate_fuse_bwd_kernel(const float* __restrict__ d_out,
                                                                 const float* __restrict__ g,
                                                                 const float* __restrict__ t,
                                                                 const float* __restrict__ im, float* __restrict__ d_gl,
                                                                 float* __restrict__ d_t, float* __restrict__ d_i, int D) {
  __shared__ float sm[8];
  const long long b = blockIdx.x;
  float dot = 0.f;
  for (int c = threadIdx.x; c < D; c += kThreads) {
    const long long i = b * D + c;
    dot += d_out[i] * (t[i] - im[i]) * g[i];
  }
  dot = block_reduce(dot, sm, false);
  for (int c = threadIdx.x; c < D; c += kThreads) {
    const long long i = b * D + c;
    const float dg = d_out[i] * (t[i] - im[i]);
    d_gl[i] = g[i] * (dg - dot);
    d_t[i] = d_out[i] * g[i];
    d_i[i] = d_out[i] * (1.f - g[i]);
  }
}


// ------------------------------------------------------------------ JS divergence of two row-softmaxes: block per row
// XModules.py:32-41.  Two passes over the row for the softmax statistics of p and q, one for the terms.
struct RowStats { float mp, mq, ip, iq; };   // row maxima and 1 / sum exp

__device__ __forceinline__ RowStats js_row_stats(const float* pr, const float* qr, int cols, int get_softmax, float* sm) {
  RowStats s{0.f, 0.f, 1.f, 1.f};
  if (!get_softmax) return s;
  float mp = -INFINITY, mq = -INFINITY;
  for (int c = threadIdx.x; c < cols; c += kThreads) {
    mp = fmaxf(mp, pr[c]);
    mq = fmaxf(mq, qr[c]);
  }
  mp = block_reduce(mp, sm, true);
  mq = block_reduce(mq, sm, true);
  float sp = 0.f, sq = 0.f;
  for (int c = threadIdx.x; c < cols; c += kThreads) {
    sp += expf(pr[c] - mp);
    sq += expf(qr[c] - mq);
  }
  sp = block_reduce(sp, sm, false);
  sq = block_reduce(sq, sm, false);
  s.mp = mp; s.mq = mq; s.ip = 1.f / sp; s.iq = 1.f / sq;
  return s;
}

__global__ void __launch_bounds__(kThreads) js_div_fwd_kernel(const float* __restrict__ p, const float* __restrict__ q,
                                                              int cols, int get_softmax, float scale,
                                                              float* __restrict__ loss) {
  __shared__ float sm[8];
  const float* pr = p + (long long)blockIdx.x * cols;
  const float* qr = q + (long long)blockIdx.x * cols;
  const RowStats s = js_row_stats(pr, qr, cols, get_softmax, sm);
  float acc = 0.f;
  for (int c = threadIdx.x; c < cols; c += kThreads) {
    const float P = get_softmax ? expf(pr[c] - s.mp) * s.ip : pr[c];
    const float Q = get_softmax ? expf(qr[c] - s.mq) * s.iq : qr[c];
    const float lm = logf(0.5f * (P + Q));
    if (P > 0.f) acc += P * (logf(P) - lm);            // xlogy semantics of KLDivLoss: 0 log 0 = 0
    if (Q > 0.f) acc += Q * (logf(Q) - lm);
  }
  acc = block_reduce(acc, sm, false);
  if (threadIdx.x == 0) atomicAdd(loss, acc * scale);
}

__global__ void __launch_bounds__(kThreads) js_div_bwd_kernel(const float* __restrict__ p, const float* __restrict__ q,
                                                              int cols, int get_softmax, float scale,
                                                              const float* __restrict__ d_loss, float* __restrict__ dp,
                                                              float* __restrict__ dq) {
  __shared__ float sm[8];
  const long long row = blockIdx.x;
  const float* pr = p + row * cols;
  const float* qr = q + row * cols;
  const RowStats s = js_row_stats(pr, qr, cols, get_softmax, sm);
  const float up = d_loss[0] * scale;
  // d f / d P_j = (log P_j - log M_j) / 2 (the +1 of x log x cancels against the M terms); same for Q
  float dotp = 0.f, dotq = 0.f;
  for (int c = threadIdx.x; c < cols; c += kThreads) {
    const float P = get_softmax ? expf(pr[c] - s.mp) * s.ip : pr[c];
    const float Q = get_softmax ? expf(qr[c] - s.mq) * s.iq : qr[c];
    const float lm = logf(0.5f * (P + Q));
    const float gp = P > 0.f ? logf(P) - lm : 0.f;
    const float gq = Q > 0.f ? logf(Q) - lm : 0.f;
    dotp += P * gp;
    dotq += Q * gq;
  }
  if (get_softmax) {
    dotp = block_reduce(dotp, sm, false);
    dotq = block_reduce(dotq, sm, false);
  }
  for (int c = threadIdx.x; c < cols; c += kThreads) {
    const float P = get_softmax ? expf(pr[c] - s.mp) * s.ip : pr[c];
    const float Q = get_softmax ? expf(qr[c] - s.mq) * s.iq : qr[c];
    const float lm = logf(0.5f * (P + Q));
    const float gp = P > 0.f ? logf(P) - lm : 0.f;
    const float gq = Q > 0.f ? logf(Q) - lm : 0.f;
    // through the softmax: dp_j = P_j (g_j - sum_k P_k g_k)
    dp[row * cols + c] = up * (get_softmax ? P * (gp - dotp) : gp);
    dq[row * cols + c] = up * (get_softmax ? Q * (gq - dotq) : gq);
  }
}

// ------------------------------------------------------------------ Block fusion core: one warp per (sample, chunk)
// XModules.py:538-543.  Lanes own the chunk positions s = lane + 32 i (S <= 128).
template <typename T>
__global__ void __launch_bounds__(kThreads) block_merge_fwd_kernel(const T* __restrict__ m0, const T* __restrict__ m1,
                                                                   long long units, int R, int S, T* __restrict__ z,
                                                                   float* __restrict__ r, float* __restrict__ inv_norm) {
  const int lane = threadIdx.x & 31;
  const long long u = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);   // = b * C + c
  if (u >= units) return;
  const T* a = m0 + u * R * S;
  const T* b = m1 + u * R * S;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k = 0; k < R; ++k) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int s = lane + 32 * i;
      if (s < S) acc[i] = fmaf(Elem<T>::ld(a + k * S + s), Elem<T>::ld(b + k * S + s), acc[i]);
    }
  }
  float zs[4], n2 = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    zs[i] = copysignf(sqrtf(fabsf(acc[i])), acc[i]);
    if (lane + 32 * i < S) n2 = fmaf(zs[i], zs[i], n2);
  }
  n2 = warp_sum(n2);
  const float inv = 1.f / fmaxf(sqrtf(n2), 1e-12f);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int s = lane + 32 * i;
    if (s < S) {
      Elem<T>::st(z + u * S + s, zs[i] * inv);
      r[u * S + s] = acc[i];
    }
  }
  if (lane == 0) inv_norm[u] = inv;
}

template <typename T>
__global__ void __launch_bounds__(kThreads) block_merge_bwd_kernel(const T* __restrict__ dz, const T* __restrict__ m0,
                                                                   const T* __restrict__ m1, const float* __restrict__ r,
                                                                   const float* __restrict__ inv_norm, long long units,
                                                                   int R, int S, T* __restrict__ dm0,
                                                                   T* __restrict__ dm1) {
  const int lane = threadIdx.x & 31;
  const long long u = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (u >= units) return;
  const float inv = inv_norm[u];
  float rv[4], zn[4], g[4], dot = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int s = lane + 32 * i;
    rv[i] = s < S ? r[u * S + s] : 0.f;
    g[i] = s < S ? Elem<T>::ld(dz + u * S + s) : 0.f;
    zn[i] = copysignf(sqrtf(fabsf(rv[i])), rv[i]) * inv;
    dot = fmaf(zn[i], g[i], dot);
  }
  dot = warp_sum(dot);
  float dr[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float dzs = (g[i] - zn[i] * dot) * inv;               // through z = zs / ||zs||
    const float sq = sqrtf(fabsf(rv[i]));
    dr[i] = sq > 0.f ? dzs * 0.5f / sq : 0.f;                    // through zs = sign(r) sqrt|r|
  }
  const T* a = m0 + u * R * S;
  const T* b = m1 + u * R * S;
  for (int k = 0; k < R; ++k) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int s = lane + 32 * i;
      if (s < S) {
        Elem<T>::st(dm0 + u * R * S + k * S + s, dr[i] * Elem<T>::ld(b + k * S + s));
        Elem<T>::st(dm1 + u * R * S + k * S + s, dr[i] * Elem<T>::ld(a + k * S + s));
      }
    }
  }
}
}  // namespace

// ======================================================================== C ABI
extern "C" {

int d2r_cast(const void* x, int32_t x_dtype, void* y, int32_t y_dtype, int64_t n, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  if (n <= 0) return D2R_OK;
  D2R_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0, "cast: pointers must be 16-byte aligned");
  const unsigned grid = blocks_for(n / 8 + 1, kThreads);
  if (x_dtype == D2R_F32 && y_dtype == D2R_BF16)
    cast_kernel<float, __nv_bfloat16><<<grid, kThreads, 0, st>>>((const float*)x, (__nv_bfloat16*)y, n);
  else if (x_dtype == D2R_BF16 && y_dtype == D2R_F32)
    cast_kernel<__nv_bfloat16, float><<<grid, kThreads, 0, st>>>((const __nv_bfloat16*)x, (float*)y, n);
  else if (x_dtype == D2R_F32 && y_dtype == D2R_F32)
    cast_kernel<float, float><<<grid, kThreads, 0, st>>>((const float*)x, (float*)y, n);
  else if (x_dtype == D2R_BF16 && y_dtype == D2R_BF16)
    cast_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, kThreads, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, n);
  else
    return set_error(D2R_ERR_ARG, "cast: bad dtypes %d -> %d", x_dtype, y_dtype);
  count_launch();
  return check_launch("cast_kernel");
}

int d2r_axpby(const void* x, const void* z, int32_t dtype, float a, float b, void* y, int64_t n, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  if (n <= 0) return D2R_OK;
  const unsigned grid = blocks_for(n / 8 + 1, kThreads);
  D2R_DISPATCH_DTYPE(dtype, T, axpby_kernel<T><<<grid, kThreads, 0, st>>>((const T*)x, (const T*)z, a, b, (T*)y, n));
  count_launch();
  return check_launch("axpby_kernel");
}

int d2r_bias_act_bwd(const void* dy, const void* y, int32_t dtype, int32_t act, void* dz, float* db, int64_t rows,
                     int32_t cols, int64_t ld, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  D2R_CHECK_ARG(cols % 8 == 0 && ld % 8 == 0, "bias_act_bwd: cols/ld must be multiples of 8");
  D2R_CHECK_ARG(act == D2R_ACT_NONE || y != nullptr, "bias_act_bwd: activation needs y");
  if (rows <= 0) return D2R_OK;
  // ~4 resident blocks per SM in ONE wave: a block streams a tall slab (>= 64 rows) so that its shared-memory
  // reduction and its 256 atomics are amortised over many rows instead of paid once per 64 rows
  const unsigned gx = (unsigned)((cols + 255) / 256);
  const long long want_y = (4LL * num_sms() + gx - 1) / gx;
  long long rpb = (rows + want_y - 1) / want_y;
  rpb = (rpb + kRowsPerBlock - 1) / kRowsPerBlock * kRowsPerBlock;
  dim3 grid(gx, (unsigned)((rows + rpb - 1) / rpb));
  if (act == D2R_ACT_NONE) {
    D2R_DISPATCH_DTYPE(dtype, T, bias_act_bwd_kernel<T, false><<<grid, kThreads, 0, st>>>(
                                     (const T*)dy, (const T*)y, act, (T*)dz, db, rows, cols, ld, (int)rpb));
  } else {
    D2R_DISPATCH_DTYPE(dtype, T, bias_act_bwd_kernel<T, true><<<grid, kThreads, 0, st>>>(
                                     (const T*)dy, (const T*)y, act, (T*)dz, db, rows, cols, ld, (int)rpb));
  }
  count_launch();
  return check_launch("bias_act_bwd_kernel");
}

#define D2R_SOFTMAX_FWD(TX, TY)                                                                              \
  do {                                                                                                       \
    if (cols <= 64)                                                                                          \
      softmax_fwd_kernel<TX, TY, 2><<<grid, kThreads, 0, st>>>((const TX*)x, ldx, (TY*)y, ldy, rows, cols, scale); \
    else if (cols <= 256)                                                                                    \
      softmax_fwd_kernel<TX, TY, 8><<<grid, kThreads, 0, st>>>((const TX*)x, ldx, (TY*)y, ldy, rows, cols, scale); \
    else                                                                                                     \
      softmax_fwd_kernel<TX, TY, 32><<<grid, kThreads, 0, st>>>((const TX*)x, ldx, (TY*)y, ldy, rows, cols, scale); \
  } while (0)

int d2r_softmax_fwd(const void* x, int32_t x_dtype, int64_t ldx, void* y, int32_t y_dtype, int64_t ldy, int64_t rows,
                    int32_t cols, float scale, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  D2R_CHECK_ARG(cols >= 1 && cols <= 1024, "softmax: cols %d outside [1,1024]", cols);
  if (rows <= 0) return D2R_OK;
  const unsigned grid = (unsigned)((rows + 7) / 8);
  if (x_dtype == D2R_F32 && y_dtype == D2R_BF16) D2R_SOFTMAX_FWD(float, __nv_bfloat16);
  else if (x_dtype == D2R_F32 && y_dtype == D2R_F32) D2R_SOFTMAX_FWD(float, float);
  else if (x_dtype == D2R_BF16 && y_dtype == D2R_BF16) D2R_SOFTMAX_FWD(__nv_bfloat16, __nv_bfloat16);
  else return set_error(D2R_ERR_ARG, "softmax_fwd: unsupported dtypes %d -> %d", x_dtype, y_dtype);
  count_launch();
  return check_launch("softmax_fwd_kernel");
}

#define D2R_SOFTMAX_BWD(TY, TD, TX)                                                                    \
  do {                                                                                                 \
    if (cols <= 64)                                                                                    \
      softmax_bwd_kernel<TY, TD, TX, 2><<<grid, kThreads, 0, st>>>((const TY*)y, ldy, (const TD*)dy, lddy, \
                                                                   (TX*)dx, lddx, rows, cols, scale);  \
    else if (cols <= 256)                                                                              \
      softmax_bwd_kernel<TY, TD, TX, 8><<<grid, kThreads, 0, st>>>((const TY*)y, ldy, (const TD*)dy, lddy, \
                                                                   (TX*)dx, lddx, rows, cols, scale);  \
    else                                                                                               \
      softmax_bwd_kernel<TY, TD, TX, 32><<<grid, kThreads, 0, st>>>((const TY*)y, ldy, (const TD*)dy, lddy, \
                                                                    (TX*)dx, lddx, rows, cols, scale); \
  } while (0)

int d2r_softmax_bwd(const void* y, int32_t y_dtype, int64_t ldy, const void* dy, int32_t dy_dtype, int64_t lddy,
                    void* dx, int32_t dx_dtype, int64_t lddx, int64_t rows, int32_t cols, float scale, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  D2R_CHECK_ARG(cols >= 1 && cols <= 1024, "softmax: cols %d outside [1,1024]", cols);
  if (rows <= 0) return D2R_OK;
  const unsigned grid = (unsigned)((rows + 7) / 8);
  if (y_dtype == D2R_BF16 && dy_dtype == D2R_F32 && dx_dtype == D2R_BF16)
    D2R_SOFTMAX_BWD(__nv_bfloat16, float, __nv_bfloat16);
  else if (y_dtype == D2R_F32 && dy_dtype == D2R_F32 && dx_dtype == D2R_F32)
    D2R_SOFTMAX_BWD(float, float, float);
  else if (y_dtype == D2R_BF16 && dy_dtype == D2R_BF16 && dx_dtype == D2R_BF16)
    D2R_SOFTMAX_BWD(__nv_bfloat16, __nv_bfloat16, __nv_bfloat16);
  else
    return set_error(D2R_ERR_ARG, "softmax_bwd: unsupported dtypes y=%d dy=%d dx=%d", y_dtype, dy_dtype, dx_dtype);
  count_launch();
  return check_launch("softmax_bwd_kernel");
}

int d2r_l2norm_fwd(const void* x, int32_t dtype, void* y, float* rnorm, int64_t rows, int32_t cols, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  D2R_CHECK_ARG(cols % 8 == 0, "l2norm: cols must be a multiple of 8");
  if (rows <= 0) return D2R_OK;
  const unsigned grid = (unsigned)((rows + 7) / 8);
  D2R_DISPATCH_DTYPE(dtype, T, l2norm_fwd_kernel<T><<<grid, kThreads, 0, st>>>((const T*)x, (T*)y, rnorm, rows, cols));
  count_launch();
  return check_launch("l2norm_fwd_kernel");
}

int d2r_l2norm_bwd(const void* y, const void* dy, int32_t dtype, const float* rnorm, void* dx, int64_t rows,
                   int32_t cols, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  D2R_CHECK_ARG(cols % 8 == 0, "l2norm: cols must be a multiple of 8");
  if (rows <= 0) return D2R_OK;
  const unsigned grid = (unsigned)((rows + 7) / 8);
  D2R_DISPATCH_DTYPE(dtype, T,
                     l2norm_bwd_kernel<T><<<grid, kThreads, 0, st>>>((const T*)y, (const T*)dy, rnorm, (T*)dx, rows, cols));
  count_launch();
  return check_launch("l2norm_bwd_kernel");
}

int d2r_film_fwd(const void* x, const void* st_, int32_t dtype, void* m, int64_t rows, int32_t cols, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  D2R_CHECK_ARG(cols % 8 == 0, "film: cols must be a multiple of 8");
  if (rows <= 0) return D2R_OK;
  const unsigned grid = blocks_for(rows * (cols / 8), kThreads);
  D2R_DISPATCH_DTYPE(dtype, T, film_fwd_kernel<T><<<grid, kThreads, 0, st>>>((const T*)x, (const T*)st_, (T*)m, rows, cols));
  count_launch();
  return check_launch("film_fwd_kernel");
}

int d2r_film_bwd(const void* dm, const void* x, const void* st_, const void* add, int32_t dtype, void* dx, void* d_st,
                 int64_t rows, int32_t cols, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  D2R_CHECK_ARG(cols % 8 == 0, "film: cols must be a multiple of 8");
  if (rows <= 0) return D2R_OK;
  const unsigned grid = blocks_for(rows * (cols / 8), kThreads);
  D2R_DISPATCH_DTYPE(dtype, T,
                     film_bwd_kernel<T><<<grid, kThreads, 0, st>>>((const T*)dm, (const T*)x, (const T*)st_, (const T*)add, (T*)dx,
                                                                  (T*)d_st, rows, cols));
  count_launch();
  return check_launch("film_bwd_kernel");
}

int d2r_mul(const void* x, const void* z, int32_t dtype, float alpha, void* y, int64_t n, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  if (n <= 0) return D2R_OK;
  const unsigned grid = blocks_for(n, kThreads);
  D2R_DISPATCH_DTYPE(dtype, T, mul_kernel<T><<<grid, kThreads, 0, st>>>((const T*)x, (const T*)z, alpha, (T*)y, n));
  count_launch();
  return check_launch("mul_kernel");
}

int d2r_sqdiff_bwd(const void* dsq, const void* d, const void* add, int32_t dtype, void* g, void* gx, int64_t n,
                   void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  if (n <= 0) return D2R_OK;
  const unsigned grid = blocks_for(n / 8 + 1, kThreads);
  D2R_DISPATCH_DTYPE(dtype, T, sqdiff_bwd_kernel<T><<<grid, kThreads, 0, st>>>((const T*)dsq, (const T*)d, (const T*)add, (T*)g, (T*)gx, n));
  count_launch();
  return check_launch("sqdiff_bwd_kernel");
}

int d2r_gate_fuse_fwd(const float* gl, const float* t, const float* i, float* g, float* out, int64_t B, int32_t D,
                      void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  if (B <= 0) return D2R_OK;
  gate_fuse_fwd_kernel<<<(unsigned)B, kThreads, 0, st>>>(gl, t, i, g, out, D);
  count_launch();
  return check_launch("gate_fuse_fwd_kernel");
}

int d2r_gate_fuse_bwd(const float* d_out, const float* g, const float* t, const float* i, float* d_gl, float* d_t,
                      float* d_i, int64_t B, int32_t D, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  if (B <= 0) return D2R_OK;
  gate_fuse_bwd_kernel<<<(unsigned)B, kThreads, 0, st>>>(d_out, g, t, i, d_gl, d_t, d_i, D);
  count_launch();
  return check_launch("gate_fuse_bwd_kernel");
}

int d2r_js_div_fwd(const float* p, const float* q, int64_t rows, int32_t cols, int32_t get_softmax, float* loss,
                   void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  D2R_CHECK_ARG(p && q && loss && cols >= 1 && rows >= 0, "js_div_fwd: bad arguments");
  if (rows == 0) return D2R_OK;
  js_div_fwd_kernel<<<(unsigned)rows, kThreads, 0, st>>>(p, q, cols, get_softmax, 0.5f / (float)rows, loss);
  count_launch();
  return check_launch("js_div_fwd_kernel");
}

int d2r_js_div_bwd(const float* p, const float* q, int64_t rows, int32_t cols, int32_t get_softmax,
                   const float* d_loss, float* dp, float* dq, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  D2R_CHECK_ARG(p && q && d_loss && dp && dq && cols >= 1 && rows >= 0, "js_div_bwd: bad arguments");
  if (rows == 0) return D2R_OK;
  js_div_bwd_kernel<<<(unsigned)rows, kThreads, 0, st>>>(p, q, cols, get_softmax, 0.5f / (float)rows, d_loss, dp, dq);
  count_launch();
  return check_launch("js_div_bwd_kernel");
}

int d2r_block_merge_fwd(const void* m0, const void* m1, int32_t dtype, int64_t B, int32_t C, int32_t R, int32_t S,
                        void* z, float* r, float* inv_norm, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  D2R_CHECK_ARG(m0 && m1 && z && r && inv_norm && C >= 1 && R >= 1 && S >= 1 && S <= 128 && B >= 0,
                "block_merge_fwd: bad arguments (chunk size must be <= 128)");
  const long long units = (long long)B * C;
  if (units == 0) return D2R_OK;
  const unsigned grid = (unsigned)((units + kThreads / 32 - 1) / (kThreads / 32));
  D2R_DISPATCH_DTYPE(dtype, T, block_merge_fwd_kernel<T><<<grid, kThreads, 0, st>>>((const T*)m0, (const T*)m1, units,
                                                                                     R, S, (T*)z, r, inv_norm));
  count_launch();
  return check_launch("block_merge_fwd_kernel");
}

int d2r_block_merge_bwd(const void* dz, const void* m0, const void* m1, const float* r, const float* inv_norm,
                        int32_t dtype, int64_t B, int32_t C, int32_t R, int32_t S, void* dm0, void* dm1,
                        void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  D2R_CHECK_ARG(dz && m0 && m1 && r && inv_norm && dm0 && dm1 && C >= 1 && R >= 1 && S >= 1 && S <= 128 && B >= 0,
                "block_merge_bwd: bad arguments (chunk size must be <= 128)");
  const long long units = (long long)B * C;
  if (units == 0) return D2R_OK;
  const unsigned grid = (unsigned)((units + kThreads / 32 - 1) / (kThreads / 32));
  D2R_DISPATCH_DTYPE(dtype, T, block_merge_bwd_kernel<T><<<grid, kThreads, 0, st>>>(
                                   (const T*)dz, (const T*)m0, (const T*)m1, r, inv_norm, units, R, S, (T*)dm0,
                                   (T*)dm1));
  count_launch();
  return check_launch("block_merge_bwd_kernel");
}

}  // extern "C"
}  // namespace d2r
