// Attention filtration of the GLAC cell (reference models/XModules.py:366-394, used at
// Cells.py:163): sim_emb S = [global ; local] (B, L+1, D)
//   logit = w.S + b  ->  BatchNorm1d(1) over all B*(L+1) scalars  ->  sigmoid  ->  l1norm over l
//   out   = l2norm(sum_l a_l S_l)                                              (B, D)
// Training mode uses batch statistics and updates the running statistics (momentum 0.1,
// unbiased variance); eval mode uses the running statistics.  The batch statistic couples all
// samples of the (per-rank) batch, exactly like the reference (no SyncBN).
#include "common.cuh"
#include "ptx.cuh"

namespace d2r {
namespace {

constexpr int kThreads = 256;
constexpr float kBnEps = 1e-5f;
constexpr float kBnMomentum = 0.1f;

template <typename T>
__device__ __forceinline__ const T* saf_row(const T* sg, const T* sl, long long b, long long l, long long L,
                                            long long D) {
  return l == 0 ? sg + b * D : sl + (b * L + (l - 1)) * D;
}

// warp per (b,l) row: logit = w . S[b,l,:] + bias
template <typename T>
__global__ void __launch_bounds__(kThreads) saf_logit_kernel(const T* __restrict__ sg, const T* __restrict__ sl,
                                                             const float* __restrict__ w,
                                                             const float* __restrict__ bias, long long B, long long L,
                                                             long long D, float* __restrict__ logits) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (row >= B * (L + 1)) return;
  const long long b = row / (L + 1), l = row % (L + 1);
  const T* s = saf_row(sg, sl, b, l, L, D);
  float acc = 0.f;
  for (long long c = lane * 8; c < D; c += 256) {
    float v[8], wv[8];
    load8(s + c, v);
    load8(w + c, wv);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc = fmaf(v[j], wv[j], acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) logits[row] = acc + bias[0];
}

__device__ __forceinline__ float block_sum(float v, float* sm) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  float r = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r += sm[w];
  return r;
}

// single block: statistics over all N = B*(L+1) logits; stats = {mean, invstd}
__global__ void __launch_bounds__(1024) saf_stats_kernel(const float* __restrict__ logits, long long N, int training,
                                                         float* running_mean, float* running_var,
                                                         long long* num_batches_tracked, float* __restrict__ stats) {
  __shared__ float sm[32];
  if (!training) {
    if (threadIdx.x == 0) {
      stats[0] = running_mean[0];
      stats[1] = rsqrtf(running_var[0] + kBnEps);
    }
    return;
  }
  float s = 0.f;
  for (long long i = threadIdx.x; i < N; i += blockDim.x) s += logits[i];
  const float mean = block_sum(s, sm) / (float)N;
  float v = 0.f;
  for (long long i = threadIdx.x; i < N; i += blockDim.x) {
    const float d = logits[i] - mean;
    v += d * d;
  }
  const float var = block_sum(v, sm) / (float)N;   // biased, used for normalisation
  if (threadIdx.x == 0) {
    stats[0] = mean;
    stats[1] = rsqrtf(var + kBnEps);
    const float unbiased = N > 1 ? var * (float)N / (float)(N - 1) : var;
    running_mean[0] = (1.f - kBnMomentum) * running_mean[0] + kBnMomentum * mean;
    running_var[0] = (1.f - kBnMomentum) * running_var[0] + kBnMomentum * unbiased;
    if (num_batches_tracked) num_batches_tracked[0] += 1;
  }
}

// block per sample: a = l1norm(sigmoid(bn(logit))); saf = sum_l a_l S_l; out = l2norm(saf)
// kOutThreads = 3 row phases x 96 column groups of 8 (16-byte loads, four rows in flight per thread); round 1 read
// the [L+1, D] slab with scalar 2-byte loads, one row at a time (0.11 of the HBM peak).
constexpr int kOutGroups = 96;
constexpr int kOutPhases = 3;
constexpr int kOutThreads = kOutGroups * kOutPhases;

template <typename T>
__global__ void __launch_bounds__(kOutThreads) saf_out_kernel(const T* __restrict__ sg, const T* __restrict__ sl,
                                                              const float* __restrict__ logits,
                                                              const float* __restrict__ stats,
                                                              const float* __restrict__ bn_w,
                                                              const float* __restrict__ bn_b, long long L, long long D,
                                                              float* __restrict__ attn, float* __restrict__ rnorm,
                                                              float* __restrict__ out) {
  extern __shared__ float s_dyn[];    // attn [L+1] (padded to a multiple of 4), then partial sums [kOutPhases][D]
  float* s_attn = s_dyn;
  float* s_part = s_dyn + ((L + 1 + 3) / 4) * 4;
  __shared__ float sm[32];
  const long long b = blockIdx.x;
  const float mean = stats[0], invstd = stats[1], gw = bn_w[0], gb = bn_b[0];
  float part = 0.f;
  for (long long l = threadIdx.x; l <= L; l += kOutThreads) {
    const float y = (logits[b * (L + 1) + l] - mean) * invstd * gw + gb;
    const float sg_ = 1.f / (1.f + __expf(-y));
    s_attn[l] = sg_;
    part += sg_;
  }
  const float tot = block_sum(part, sm);
  const float inv = 1.f / (tot + 1e-8f);
  for (long long l = threadIdx.x; l <= L; l += kOutThreads) {
    const float a = s_attn[l] * inv;
    s_attn[l] = a;
    attn[b * (L + 1) + l] = a;
  }
  __syncthreads();
  const int rp = threadIdx.x / kOutGroups;
  const long long G = D / 8;
  for (long long g = threadIdx.x % kOutGroups; g < G; g += kOutGroups) {
    const long long c = g * 8;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (long long l = rp; l <= L; l += 4 * kOutPhases) {
      float v[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long lu = l + u * kOutPhases;
        if (lu <= L) {
          load8(saf_row(sg, sl, b, lu, L, D) + c, v[u]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[u][j] = 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long lu = l + u * kOutPhases;
        const float a = lu <= L ? s_attn[lu] : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(a, v[u][j], acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s_part[rp * D + c + j] = acc[j];
  }
  __syncthreads();
  float ss = 0.f;
  for (long long c = threadIdx.x; c < D; c += kOutThreads) {
    float a = 0.f;
#pragma unroll
    for (int r = 0; r < kOutPhases; ++r) a += s_part[r * D + c];
    s_part[c] = a;                    // (phase 0's slot of column c is only read by this thread)
    ss += a * a;
  }
  ss = block_sum(ss, sm);
  const float r = 1.f / (sqrtf(ss) + 1e-8f);
  if (threadIdx.x == 0) rnorm[b] = r;
  for (long long c = threadIdx.x; c < D; c += kOutThreads) out[b * D + c] = s_part[c] * r;
}

// ---------------------------------------------------------------- backward
// block per sample: d_saf (l2norm bwd), d_a[l] = d_saf . S_l, l1norm+sigmoid bwd -> d_y[l]
template <typename T>
__global__ void __launch_bounds__(kThreads) saf_bwd_a_kernel(const T* __restrict__ sg, const T* __restrict__ sl,
                                                             const float* __restrict__ d_out,
                                                             const float* __restrict__ out,
                                                             const float* __restrict__ rnorm,
                                                             const float* __restrict__ logits,
                                                             const float* __restrict__ attn,
                                                             const float* __restrict__ stats,
                                                             const float* __restrict__ bn_w,
                                                             const float* __restrict__ bn_b, long long L, long long D,
                                                             float* __restrict__ d_saf, float* __restrict__ d_y) {
  extern __shared__ float sh[];   // d_saf[D] then d_a[L+1]
  float* s_dsaf = sh;
  float* s_da = sh + D;
  __shared__ float sm[8];
  const long long b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float dot = 0.f;
  for (long long c = threadIdx.x; c < D; c += kThreads) dot += d_out[b * D + c] * out[b * D + c];
  dot = block_sum(dot, sm);
  const float r = rnorm[b];
  const float n = fmaxf(1.f / r - 1e-8f, 1e-30f);
  for (long long c = threadIdx.x; c < D; c += kThreads) {
    const float v = r * d_out[b * D + c] - out[b * D + c] * dot / n;
    s_dsaf[c] = v;
    d_saf[b * D + c] = v;
  }
  __syncthreads();
  // four rows per warp and iteration: their loads are all in flight before the first dot product starts
  for (long long l = warp; l <= L; l += 4 * (kThreads / 32)) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long c = lane * 8; c < D; c += 256) {
      float v[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long lu = l + u * (kThreads / 32);
        if (lu <= L) {
          load8(saf_row(sg, sl, b, lu, L, D) + c, v[u]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[u][j] = 0.f;
        }
      }
      float g[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = s_dsaf[c + j];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[u] = fmaf(v[u][j], g[j], acc[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float a = warp_sum(acc[u]);
      const long long lu = l + u * (kThreads / 32);
      if (lane == 0 && lu <= L) s_da[lu] = a;
    }
  }
  __syncthreads();
  // a = s / (T + eps):  d_s = (d_a - sum_l d_a a) / (T + eps);   T + eps = s / a (any l)
  const float mean = stats[0], invstd = stats[1], gw = bn_w[0], gb = bn_b[0];
  float part = 0.f, tpart = 0.f;
  for (long long l = threadIdx.x; l <= L; l += kThreads) {
    part += s_da[l] * attn[b * (L + 1) + l];
    const float y = (logits[b * (L + 1) + l] - mean) * invstd * gw + gb;
    tpart += 1.f / (1.f + __expf(-y));
  }
  const float da_a = block_sum(part, sm);
  const float tot = block_sum(tpart, sm) + 1e-8f;
  for (long long l = threadIdx.x; l <= L; l += kThreads) {
    const float y = (logits[b * (L + 1) + l] - mean) * invstd * gw + gb;
    const float s = 1.f / (1.f + __expf(-y));
    const float ds = (s_da[l] - da_a) / tot;
    d_y[b * (L + 1) + l] = ds * s * (1.f - s);
  }
}

// single block: BN backward reductions; red = {sum d_y, sum d_y xhat}; accumulates d_bn_w/d_bn_b
__global__ void __launch_bounds__(1024) saf_bwd_bn_kernel(const float* __restrict__ d_y,
                                                          const float* __restrict__ logits,
                                                          const float* __restrict__ stats, long long N,
                                                          float* __restrict__ red, float* d_bn_w, float* d_bn_b) {
  __shared__ float sm[32];
  const float mean = stats[0], invstd = stats[1];
  float s1 = 0.f, s2 = 0.f;
  for (long long i = threadIdx.x; i < N; i += blockDim.x) {
    const float g = d_y[i];
    s1 += g;
    s2 += g * (logits[i] - mean) * invstd;
  }
  s1 = block_sum(s1, sm);
  s2 = block_sum(s2, sm);
  if (threadIdx.x == 0) {
    red[0] = s1;
    red[1] = s2;
    if (d_bn_b) d_bn_b[0] += s1;
    if (d_bn_w) d_bn_w[0] += s2;
  }
}

// warp per (b,l) row: d_logit, dS = a d_saf + d_logit w; dw += d_logit S; dbias += d_logit
template <typename T>
__global__ void __launch_bounds__(kThreads) saf_bwd_s_kernel(const T* __restrict__ sg, const T* __restrict__ sl,
                                                             const float* __restrict__ w,
                                                             const float* __restrict__ logits,
                                                             const float* __restrict__ attn,
                                                             const float* __restrict__ stats,
                                                             const float* __restrict__ bn_w,
                                                             const float* __restrict__ d_saf,
                                                             const float* __restrict__ d_y,
                                                             const float* __restrict__ red, int training, long long B,
                                                             long long L, long long D, T* __restrict__ d_sg,
                                                             T* __restrict__ d_sl, float* __restrict__ d_w,
                                                             float* __restrict__ d_bias) {
  extern __shared__ float s_dw[];   // [D] per-block partial of d_w
  __shared__ float s_db;
  for (long long c = threadIdx.x; c < D; c += kThreads) s_dw[c] = 0.f;
  if (threadIdx.x == 0) s_db = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long N = B * (L + 1);
  const float mean = stats[0], invstd = stats[1], gw = bn_w[0];
  const float m1 = red[0] / (float)N, m2 = red[1] / (float)N;
  // each warp handles a contiguous chunk of rows so the smem atomics stay per-block
  const long long rows_per_block = (N + gridDim.x - 1) / gridDim.x;
  const long long r_begin = (long long)blockIdx.x * rows_per_block;
  const long long r_end = min(N, r_begin + rows_per_block);
  float wacc[4][8];   // this lane's columns lane*8 + 256*k (D <= 1024)
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) wacc[k][j] = 0.f;
  float dbacc = 0.f;
  for (long long row = r_begin + (threadIdx.x >> 5); row < r_end; row += kThreads / 32) {
    const long long b = row / (L + 1), l = row % (L + 1);
    const float xhat = (logits[row] - mean) * invstd;
    const float dl = training ? gw * invstd * (d_y[row] - m1 - xhat * m2) : gw * invstd * d_y[row];
    const float a = attn[row];
    const T* s = saf_row(sg, sl, b, l, L, D);
    T* ds = l == 0 ? d_sg + b * D : d_sl + (b * L + (l - 1)) * D;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long c = lane * 8 + 256 * k;
      if (c < D) {
        float v[8], wv[8], g[8], o[8];
        load8(s + c, v);
        load8(w + c, wv);
        load8(d_saf + b * D + c, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          o[j] = a * g[j] + dl * wv[j];
          wacc[k][j] = fmaf(dl, v[j], wacc[k][j]);
        }
        store8(ds + c, o);
      }
    }
    dbacc += dl;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long c = lane * 8 + 256 * k;
    if (c < D) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&s_dw[c + j], wacc[k][j]);
    }
  }
  if (lane == 0) atomicAdd(&s_db, dbacc);
  __syncthreads();
  for (long long c = threadIdx.x; c < D; c += kThreads) atomicAdd(d_w + c, s_dw[c]);
  if (threadIdx.x == 0) atomicAdd(d_bias, s_db);
}

}  // namespace

extern "C" {

int d2r_saf_fwd(const d2r_saf_args* a, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  D2R_CHECK_ARG(a && a->B > 0 && a->L > 0 && a->D % 8 == 0, "saf: bad shape");
  const long long N = a->B * (a->L + 1);
  const unsigned grid_rows = (unsigned)((N + 7) / 8);
  const size_t sh = sizeof(float) * (size_t)(((a->L + 1 + 3) / 4) * 4 + kOutPhases * a->D);
  D2R_CHECK_ARG(sh <= 40 * 1024, "saf: L + 3 D too large");
  D2R_DISPATCH_DTYPE(a->dtype, T,
                     saf_logit_kernel<T><<<grid_rows, kThreads, 0, st>>>((const T*)a->sg, (const T*)a->sl, a->w, a->bias,
                                                                         a->B, a->L, a->D, a->logits));
  saf_stats_kernel<<<1, 1024, 0, st>>>(a->logits, N, a->training, a->running_mean, a->running_var,
                                       (long long*)a->num_batches_tracked, a->stats);
  D2R_DISPATCH_DTYPE(a->dtype, T,
                     saf_out_kernel<T><<<(unsigned)a->B, kOutThreads, sh, st>>>((const T*)a->sg, (const T*)a->sl, a->logits,
                                                                             a->stats, a->bn_w, a->bn_b, a->L, a->D,
                                                                             a->attn, a->rnorm, a->out));
  count_launch(3);
  return check_launch("saf_fwd");
}

int d2r_saf_bwd(const d2r_saf_bwd_args* a, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  D2R_CHECK_ARG(a && a->scratch && a->d_out, "saf_bwd: null args");
  const d2r_saf_args* f = &a->fwd;
  const long long N = f->B * (f->L + 1);
  float* d_saf = a->scratch;                 // [B*D]
  float* d_y = d_saf + f->B * f->D;          // [N]
  float* red = d_y + N;                      // [2]
  const size_t sh_a = sizeof(float) * (size_t)(f->D + f->L + 1);
  const size_t sh_s = sizeof(float) * (size_t)f->D;
  D2R_CHECK_ARG(sh_a <= 40 * 1024, "saf_bwd: D + L too large");
  D2R_DISPATCH_DTYPE(f->dtype, T,
                     saf_bwd_a_kernel<T><<<(unsigned)f->B, kThreads, sh_a, st>>>(
                         (const T*)f->sg, (const T*)f->sl, a->d_out, f->out, f->rnorm, f->logits, f->attn, f->stats,
                         f->bn_w, f->bn_b, f->L, f->D, d_saf, d_y));
  saf_bwd_bn_kernel<<<1, 1024, 0, st>>>(d_y, f->logits, f->stats, N, red, a->d_bn_w, a->d_bn_b);
  long long blocks = (N + 63) / 64;
  if (blocks > 148 * 4) blocks = 148 * 4;
  D2R_DISPATCH_DTYPE(f->dtype, T,
                     saf_bwd_s_kernel<T><<<(unsigned)blocks, kThreads, sh_s, st>>>(
                         (const T*)f->sg, (const T*)f->sl, f->w, f->logits, f->attn, f->stats, f->bn_w, d_saf, d_y, red,
                         f->training, f->B, f->L, f->D, (T*)a->d_sg, (T*)a->d_sl, a->d_w, a->d_bias));
  count_launch(3);
  return check_launch("saf_bwd");
}

}  // extern "C"
}  // namespace d2r
