// (a) Router kernels.  Router.forward (reference models/Router.py:22-26):
//   soft_g = relu(tanh(W2 relu(W1 mean_L(x) + b1) + b2))
// followed by the cross-cell normalisation / gate of DynamicInteraction.py:50-52 (or :104-117).
//   d2r_pool_mean        HBM-bound mean over L (layer 0; later layers get the pooled mean for
//                        free from the aggregation kernel)
//   hidden layer         fp32 d2r_gemm (tiny: [B,768]x[768,768])
//   d2r_router_head_*    W2 / tanh / relu / normalise / gate for all K cells of a layer
#include "common.cuh"
#include "ptx.cuh"

namespace d2r {
namespace {

constexpr int kThreads = 256;

// grid (D/256, B, groups); lane owns 8 consecutive columns, 8 warps stride over L
template <typename T>
__global__ void __launch_bounds__(kThreads) pool_mean_kernel(d2r_ptr8 xs, long long B, long long L, long long D,
                                                             float* __restrict__ pooled) {
  __shared__ float red[8][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = blockIdx.y;
  const int g = blockIdx.z;
  const long long col = (long long)blockIdx.x * 256 + lane * 8;
  const T* x = static_cast<const T*>(xs.p[g]) + b * L * D;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (col < D) {
    // four rows (4 x 16 B) in flight per thread before the first is consumed
    for (long long l = warp; l < L; l += 32) {
      float v[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (l + 8 * u < L) {
          load8(x + (l + 8 * u) * D + col, v[u]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[u][j] = 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[u][j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = acc[j];
  __syncthreads();
  const int c = threadIdx.x;
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += red[w][c];
  const long long oc = (long long)blockIdx.x * 256 + c;
  if (oc < D) pooled[((long long)g * B + b) * D + oc] = s / (float)L;
}

// dx[b,l,:] (+)= d_pooled[b,:] / L
template <typename T>
__global__ void pool_mean_bwd_kernel(const float* __restrict__ dp, long long B, long long L, long long D,
                                     T* __restrict__ dx, int accumulate) {
  const long long vpr = D / 8;
  const long long total = B * L * vpr;
  const float inv = 1.f / (float)L;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long c = (i % vpr) * 8;
    const long long bl = i / vpr;
    const long long b = bl / L;
    float g[8];
    load8(dp + b * D + c, g);
    float v[8];
    if (accumulate) {
      load8(dx + bl * D + c, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += g[j] * inv;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = g[j] * inv;
    }
    store8(dx + bl * D + c, v);
  }
}

// block per sample b.  K cells x n_out paths dot products of length H, one warp each (round robin).
__global__ void __launch_bounds__(kThreads) router_head_fwd_kernel(const float* __restrict__ hid, d2r_ptr8 w2,
                                                                   d2r_ptr8 b2, int K, int n_out, long long B, int H,
                                                                   int final_layer, float* __restrict__ raw,
                                                                   float* __restrict__ norm, float* __restrict__ gate) {
  __shared__ float s_raw[64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = blockIdx.x;
  for (int e = warp; e < K * n_out; e += 8) {
    const int j = e / n_out, i = e % n_out;   // cell j, out path i
    const float* h = hid + ((long long)j * B + b) * H;
    const float* w = static_cast<const float*>(w2.p[j]) + (long long)i * H;
    float acc = 0.f;
    for (int c = lane * 4; c < H; c += 128) {
      const float4 hv = *reinterpret_cast<const float4*>(h + c);
      const float4 wv = __ldg(reinterpret_cast<const float4*>(w + c));
      acc += hv.x * wv.x + hv.y * wv.y + hv.z * wv.z + hv.w * wv.w;
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      const float logit = acc + static_cast<const float*>(b2.p[j])[i];
      s_raw[i * K + j] = fmaxf(tanhf(logit), 0.f);
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < n_out) {
    const int i = threadIdx.x;
    float sum = 0.f;
    for (int j = 0; j < K; ++j) sum += s_raw[i * K + j];
    for (int j = 0; j < K; ++j) {
      const float r = s_raw[i * K + j];
      raw[(b * n_out + i) * K + j] = r;
      norm[(b * n_out + i) * K + j] = final_layer ? r : r / (sum + 1e-8f);
    }
    if (!final_layer) gate[b * n_out + i] = sum < 1e-4f ? 1.f : 0.f;
  }
  if (final_layer && (int)threadIdx.x < K)
    gate[b * K + threadIdx.x] = s_raw[threadIdx.x] < (1e-4f / (float)K) ? 1.f : 0.f;
}

// Backward of the router head for all K cells of a layer, ONE launch.  grid (K, ceil(B / kRhbTile)): a block owns cell j
// and a tile of kRhbTile samples.
//   d_logit[b,i,j]  from d_norm / raw (the cross-cell normalisation couples the K cells of a (b, i) row; every block
//                   recomputes the row sums it needs -- K values)
//   d_hid[j,b,:]  = relu'(hid) * sum_i d_logit[b,i,j] W2_j[i,:]      W2_j rows are read ONCE per block, not per sample
//   dW2_j[i,:]   += sum_b d_logit[b,i,j] hid[j,b,:]                  one atomicAdd per (i, column) and block
//   db2_j[i]     += sum_b d_logit[b,i,j]
// (round 1: one block per sample with a serial loop over the cells re-reading W2 for every sample, plus a second
//  kernel for the weight gradient: 108 us per call at B=256; this one moves 2 x [K,B,H] + K x [n_out,H] floats once.)
constexpr int kRhbTile = 8;
constexpr int kRhbMaxOut = 8;

__global__ void __launch_bounds__(kThreads) router_head_bwd_kernel(const float* __restrict__ d_norm,
                                                                   const float* __restrict__ raw,
                                                                   const float* __restrict__ hid, d2r_ptr8 w2, int K,
                                                                   int n_out, long long B, int H, int final_layer,
                                                                   float* __restrict__ d_hid,
                                                                   float* __restrict__ d_logit, d2r_ptr8 d_w2,
                                                                   d2r_ptr8 d_b2) {
  __shared__ float s_dl[kRhbTile][kRhbMaxOut];
  const int j = blockIdx.x;
  const long long b0 = (long long)blockIdx.y * kRhbTile;
  const int nb = (int)min((long long)kRhbTile, B - b0);
  if (threadIdx.x < kRhbTile * kRhbMaxOut) s_dl[threadIdx.x / kRhbMaxOut][threadIdx.x % kRhbMaxOut] = 0.f;
  __syncthreads();
  if ((int)threadIdx.x < kRhbTile * n_out) {
    const int t = threadIdx.x / n_out, i = threadIdx.x % n_out;
    float dl = 0.f;
    if (t < nb) {
      const long long row = ((b0 + t) * n_out + i) * K;
      float sum = 0.f, dot = 0.f;
      for (int c = 0; c < K; ++c) {
        const float r = raw[row + c];
        sum += r;
        dot += d_norm[row + c] * r;
      }
      const float inv = 1.f / (sum + 1e-8f);
      const float r = raw[row + j];
      const float dn = d_norm[row + j];
      const float d_raw = final_layer ? dn : (dn * inv - dot * inv * inv);
      // raw = relu(tanh(z)):  d/dz = (1 - tanh^2) where tanh > 0
      dl = r > 0.f ? d_raw * (1.f - r * r) : 0.f;
      d_logit[row + j] = dl;
    }
    s_dl[t][i] = dl;
  }
  __syncthreads();
  const float* w = static_cast<const float*>(w2.p[j]);
  float* dw = static_cast<float*>(const_cast<void*>(d_w2.p[j]));
  for (int c = threadIdx.x; c < H; c += kThreads) {
    float wv[kRhbMaxOut], aw[kRhbMaxOut];
#pragma unroll
    for (int i = 0; i < kRhbMaxOut; ++i) {
      wv[i] = i < n_out ? __ldg(w + (long long)i * H + c) : 0.f;
      aw[i] = 0.f;
    }
    float hv[kRhbTile];
#pragma unroll
    for (int t = 0; t < kRhbTile; ++t) hv[t] = t < nb ? hid[((long long)j * B + b0 + t) * H + c] : 0.f;
#pragma unroll
    for (int t = 0; t < kRhbTile; ++t) {
      if (t < nb) {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < kRhbMaxOut; ++i) {
          const float dl = s_dl[t][i];          // (zero for i >= n_out)
          acc = fmaf(dl, wv[i], acc);
          aw[i] = fmaf(dl, hv[t], aw[i]);
        }
        d_hid[((long long)j * B + b0 + t) * H + c] = hv[t] > 0.f ? acc : 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < kRhbMaxOut; ++i)
      if (i < n_out) atomicAdd(dw + (long long)i * H + c, aw[i]);
  }
  if ((int)threadIdx.x < n_out) {
    float sb = 0.f;
    for (int t = 0; t < nb; ++t) sb += s_dl[t][threadIdx.x];
    atomicAdd(static_cast<float*>(const_cast<void*>(d_b2.p[j])) + threadIdx.x, sb);
  }
}

}  // namespace

extern "C" {

int d2r_pool_mean(d2r_ptr8 x, int32_t groups, int32_t x_dtype, int64_t B, int64_t L, int64_t D, float* pooled,
                  void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  D2R_CHECK_ARG(groups >= 1 && groups <= 8, "pool_mean: groups %d outside [1,8]", groups);
  D2R_CHECK_ARG(D % 8 == 0 && B > 0 && L > 0 && B <= 65535, "pool_mean: bad shape B=%lld L=%lld D=%lld",
                (long long)B, (long long)L, (long long)D);
  dim3 grid((unsigned)((D + 255) / 256), (unsigned)B, (unsigned)groups);
  D2R_DISPATCH_DTYPE(x_dtype, T, pool_mean_kernel<T><<<grid, kThreads, 0, st>>>(x, B, L, D, pooled));
  count_launch();
  return check_launch("pool_mean_kernel");
}

int d2r_pool_mean_bwd(const float* d_pooled, int64_t B, int64_t L, int64_t D, void* dx, int32_t dx_dtype,
                      int32_t accumulate, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  D2R_CHECK_ARG(D % 8 == 0 && B > 0 && L > 0, "pool_mean_bwd: bad shape");
  long long blocks = (B * L * (D / 8) + kThreads - 1) / kThreads;
  if (blocks > 148 * 32) blocks = 148 * 32;
  D2R_DISPATCH_DTYPE(dx_dtype, T,
                     pool_mean_bwd_kernel<T><<<(unsigned)blocks, kThreads, 0, st>>>(d_pooled, B, L, D, (T*)dx, accumulate));
  count_launch();
  return check_launch("pool_mean_bwd_kernel");
}

int d2r_router_head_fwd(const float* hid, d2r_ptr8 w2, d2r_ptr8 b2, int32_t K, int32_t n_out, int64_t B, int32_t H,
                        int32_t final_layer, float* raw, float* norm, float* gate, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  D2R_CHECK_ARG(K >= 1 && K <= 8 && n_out >= 1 && n_out <= 8 && H % 4 == 0 && B > 0, "router_head: bad shape");
  D2R_CHECK_ARG(!final_layer || n_out == 1, "router_head: the final layer has one out path");
  router_head_fwd_kernel<<<(unsigned)B, kThreads, 0, st>>>(hid, w2, b2, K, n_out, B, H, final_layer, raw, norm, gate);
  count_launch();
  return check_launch("router_head_fwd_kernel");
}

int d2r_router_head_bwd(const float* d_norm, const float* raw, const float* hid, d2r_ptr8 w2, int32_t K,
                        int32_t n_out, int64_t B, int32_t H, int32_t final_layer, float* d_hid, float* d_logit,
                        d2r_ptr8 d_w2, d2r_ptr8 d_b2, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  D2R_CHECK_ARG(K >= 1 && K <= 8 && n_out >= 1 && n_out <= 8 && B > 0, "router_head_bwd: bad shape");
  // d_w2 / d_b2 are ACCUMULATED into (atomics): the caller passes zero-initialised (or running-sum) buffers
  router_head_bwd_kernel<<<dim3((unsigned)K, (unsigned)((B + kRhbTile - 1) / kRhbTile)), kThreads, 0, st>>>(
      d_norm, raw, hid, w2, K, n_out, B, H, final_layer, d_hid, d_logit, d_w2, d_b2);
  count_launch();
  return check_launch("router_head_bwd_kernel");
}

}  // extern "C"
}  // namespace d2r
