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
    for (long long l = warp; l < L; l += 8) {
      float v[8];
      load8(x + l * D + col, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
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

// block per sample: d_norm -> d_logit [B,n_out,K] and d_hid[j,b,:] = relu'(hid) * sum_i d_logit_ij W2_j[i,:]
__global__ void __launch_bounds__(kThreads) router_head_bwd_kernel(const float* __restrict__ d_norm,
                                                                   const float* __restrict__ raw,
                                                                   const float* __restrict__ hid, d2r_ptr8 w2, int K,
                                                                   int n_out, long long B, int H, int final_layer,
                                                                   float* __restrict__ d_hid,
                                                                   float* __restrict__ d_logit) {
  __shared__ float s_dl[64];
  const long long b = blockIdx.x;
  if ((int)threadIdx.x < n_out) {
    const int i = threadIdx.x;
    float sum = 0.f, dot = 0.f;
    for (int j = 0; j < K; ++j) {
      const float r = raw[(b * n_out + i) * K + j];
      sum += r;
      dot += d_norm[(b * n_out + i) * K + j] * r;
    }
    const float inv = 1.f / (sum + 1e-8f);
    for (int j = 0; j < K; ++j) {
      const float r = raw[(b * n_out + i) * K + j];
      const float dn = d_norm[(b * n_out + i) * K + j];
      const float d_raw = final_layer ? dn : (dn * inv - dot * inv * inv);
      // raw = relu(tanh(z)):  d/dz = (1 - tanh^2) where tanh > 0
      const float dl = r > 0.f ? d_raw * (1.f - r * r) : 0.f;
      s_dl[i * K + j] = dl;
      d_logit[(b * n_out + i) * K + j] = dl;
    }
  }
  __syncthreads();
  for (int j = 0; j < K; ++j) {
    const float* w = static_cast<const float*>(w2.p[j]);
    for (int c = threadIdx.x; c < H; c += kThreads) {
      const long long idx = ((long long)j * B + b) * H + c;
      float acc = 0.f;
      for (int i = 0; i < n_out; ++i) acc += s_dl[i * K + j] * __ldg(w + (long long)i * H + c);
      d_hid[idx] = hid[idx] > 0.f ? acc : 0.f;
    }
  }
}

// grid (K*n_out, H/64): block = 64 columns x 4 batch groups.  dW2_j[i,:] += sum_b d_logit[b,i,j] hid[j,b,:];
// db2_j[i] += sum_b d_logit[b,i,j]
__global__ void __launch_bounds__(kThreads) router_head_wgrad_kernel(const float* __restrict__ d_logit,
                                                                     const float* __restrict__ hid, int K, int n_out,
                                                                     long long B, int H, d2r_ptr8 d_w2, d2r_ptr8 d_b2) {
  __shared__ float red[4][64];
  const int j = blockIdx.x / n_out, i = blockIdx.x % n_out;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int c = blockIdx.y * 64 + tx;
  float acc = 0.f, sb = 0.f;
  if (c < H) {
#pragma unroll 4
    for (long long b = ty; b < B; b += 4) {
      const float dl = d_logit[(b * n_out + i) * K + j];
      acc = fmaf(dl, hid[((long long)j * B + b) * H + c], acc);
      sb += dl;
    }
  }
  red[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && c < H) {
    float* dw = static_cast<float*>(const_cast<void*>(d_w2.p[j])) + (long long)i * H;
    dw[c] += red[0][tx] + red[1][tx] + red[2][tx] + red[3][tx];
  }
  __syncthreads();
  if (blockIdx.y == 0 && tx == 0) red[ty][0] = sb;
  __syncthreads();
  if (blockIdx.y == 0 && threadIdx.x == 0) {
    float* db = static_cast<float*>(const_cast<void*>(d_b2.p[j])) + i;
    *db += red[0][0] + red[1][0] + red[2][0] + red[3][0];
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
  router_head_bwd_kernel<<<(unsigned)B, kThreads, 0, st>>>(d_norm, raw, hid, w2, K, n_out, B, H, final_layer, d_hid,
                                                           d_logit);
  router_head_wgrad_kernel<<<dim3((unsigned)(K * n_out), (unsigned)((H + 63) / 64)), kThreads, 0, st>>>(d_logit, hid, K,
                                                                                                   n_out, B, H, d_w2, d_b2);
  count_launch(2);
  return check_launch("router_head_bwd_kernel");
}

}  // extern "C"
}  // namespace d2r
