// fp32 CUDA-core batched GEMM with the same contract and epilogue as the tcgen05 path.
// This is the arithmetic path for dtype = D2R_F32 (the "1e-5 relative / bit-exact argmax"
// parity mode of the spec, where TF32 tensor cores are not accurate enough) and for the tiny
// [B, 768] contractions of the routers and the global cells, whose results feed the routing
// probabilities and are therefore kept in fp32 in every mode.
#include "common.cuh"

namespace d2r {

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

struct SimtParams {
  int m, n, k, batch_inner;
  int a_mn, b_mn;
  const float* a;
  const float* b;
  void* c;
  void* c2;
  const float* bias;
  const void* residual;
  long long lda, ldb, ldc, ldr;
  long long a_so, a_si, b_so, b_si, c_so, c_si, r_so, r_si, bias_sz;
  float alpha;
  int act, epilogue, c_dtype, r_dtype, atomic, act_cols;
  int split_k, k_per_split;
};

// operand element (r, kk): K-major -> base[r*ld + kk]; MN-major -> base[kk*ld + r]
__device__ __forceinline__ float ld_op(const float* base, long long ld, int mn_major, int r, int kk, int rmax,
                                       int kmax) {
  if (r >= rmax || kk >= kmax) return 0.f;
  return mn_major ? __ldg(base + (long long)kk * ld + r) : __ldg(base + (long long)r * ld + kk);
}

__global__ void __launch_bounds__(256) gemm_simt_kernel(const SimtParams p) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  int z = blockIdx.z / p.split_k;
  const int ks = blockIdx.z % p.split_k;
  const int zi = z % p.batch_inner, zo = z / p.batch_inner;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const float* A = p.a + (long long)zo * p.a_so + (long long)zi * p.a_si;
  const float* B = p.b + (long long)zo * p.b_so + (long long)zi * p.b_si;
  const int k_begin = ks * p.k_per_split;
  const int k_end = min(p.k, k_begin + p.k_per_split);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = k_begin; k0 < k_end; k0 += TK) {
    // each thread loads 4 elements of each tile; index so that global reads are contiguous
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;   // 0..1023
      int r, kk;
      if (p.a_mn) { r = idx & 63; kk = idx >> 6; } else { kk = idx & 15; r = idx >> 4; }
      As[kk][r] = ld_op(A, p.lda, p.a_mn, m0 + r, k0 + kk, p.m, k_end);
      if (p.b_mn) { r = idx & 63; kk = idx >> 6; } else { kk = idx & 15; r = idx >> 4; }
      Bs[kk][r] = ld_op(B, p.ldb, p.b_mn, n0 + r, k0 + kk, p.n, k_end);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  const float* bias = p.bias ? p.bias + (long long)z * p.bias_sz : nullptr;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= p.m) continue;
    const long long c_off = (long long)zo * p.c_so + (long long)zi * p.c_si + (long long)row * p.ldc;
    const long long r_off = (long long)zo * p.r_so + (long long)zi * p.r_si + (long long)row * p.ldr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col >= p.n) continue;
      float v = p.alpha * acc[i][j];
      if (bias) v += __ldg(bias + col);
      float res = 0.f;
      if (p.residual)
        res = p.r_dtype == D2R_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.residual)[r_off + col])
                                    : reinterpret_cast<const float*>(p.residual)[r_off + col];
      if (p.epilogue == D2R_EPI_SQDIFF) {
        const float d = res - v;
        if (p.c_dtype == D2R_BF16) reinterpret_cast<__nv_bfloat16*>(p.c2)[c_off + col] = __float2bfloat16_rn(d);
        else reinterpret_cast<float*>(p.c2)[c_off + col] = d;
        v = d * d;
      } else {
        if (p.act_cols == 0 || col < p.act_cols) v = apply_act(v, p.act);
        v += res;
      }
      if (p.c_dtype == D2R_BF16) reinterpret_cast<__nv_bfloat16*>(p.c)[c_off + col] = __float2bfloat16_rn(v);
      else if (p.atomic) atomicAdd(reinterpret_cast<float*>(p.c) + c_off + col, v);
      else reinterpret_cast<float*>(p.c)[c_off + col] = v;
    }
  }
}

}  // namespace

int gemm_simt(const d2r_gemm_args& a, cudaStream_t stream) {
  D2R_CHECK_ARG(a.m > 0 && a.n > 0 && a.k > 0 && a.batch > 0 && a.batch_inner > 0, "gemm: empty problem");
  D2R_CHECK_ARG(a.batch % a.batch_inner == 0, "gemm: batch %d not a multiple of batch_inner %d", a.batch,
                a.batch_inner);
  D2R_CHECK_ARG(a.epilogue == D2R_EPI_STD || a.epilogue == D2R_EPI_SQDIFF,
                "gemm(fp32): the fused softmax epilogues exist on the bf16 tensor-core path only");
  D2R_CHECK_ARG(a.epilogue == D2R_EPI_STD || (a.residual && a.c2), "gemm: SQDIFF needs residual and c2");
  int split_k = a.split_k > 1 ? a.split_k : 1;
  SimtParams p;
  p.m = a.m; p.n = a.n; p.k = a.k; p.batch_inner = a.batch_inner;
  p.a_mn = a.a_mn_major; p.b_mn = a.b_mn_major;
  p.a = static_cast<const float*>(a.a); p.b = static_cast<const float*>(a.b);
  p.c = a.c; p.c2 = a.c2; p.bias = a.bias; p.residual = a.residual;
  p.lda = a.lda; p.ldb = a.ldb; p.ldc = a.ldc; p.ldr = a.ldr;
  p.a_so = a.a_so; p.a_si = a.a_si; p.b_so = a.b_so; p.b_si = a.b_si;
  p.c_so = a.c_so; p.c_si = a.c_si; p.r_so = a.r_so; p.r_si = a.r_si; p.bias_sz = a.bias_sz;
  p.alpha = a.alpha; p.act = a.act; p.epilogue = a.epilogue; p.c_dtype = a.c_dtype; p.r_dtype = a.r_dtype;
  int k_per = (a.k + split_k - 1) / split_k;
  k_per = (k_per + TK - 1) / TK * TK;
  p.k_per_split = k_per;
  p.split_k = (a.k + k_per - 1) / k_per;
  const bool atomic = a.accumulate || p.split_k > 1;
  D2R_CHECK_ARG(!atomic || a.c_dtype == D2R_F32, "gemm: accumulate/split_k need an fp32 C");
  D2R_CHECK_ARG(!atomic || (a.epilogue == D2R_EPI_STD && a.act == D2R_ACT_NONE && !a.residual),
                "gemm: accumulate/split_k support only the plain epilogue");
  p.atomic = atomic ? 1 : 0;
  p.act_cols = a.act_cols;
  if (p.split_k > 1 && !a.accumulate) {
    D2R_CHECK_ARG(a.batch == 1 && a.ldc == a.n, "gemm: split_k without accumulate needs a dense, unbatched C");
    D2R_CUDA_OK(cudaMemsetAsync(a.c, 0, sizeof(float) * (size_t)a.m * a.n, stream));
  }
  const long long gz = 1LL * a.batch * p.split_k;
  D2R_CHECK_ARG(gz <= 65535, "gemm(fp32): batch*split_k %lld exceeds grid.z", gz);
  dim3 grid((a.n + TN - 1) / TN, (a.m + TM - 1) / TM, (unsigned)gz);
  D2R_CHECK_ARG(grid.y <= 65535, "gemm(fp32): too many row tiles");
  gemm_simt_kernel<<<grid, 256, 0, stream>>>(p);
  count_launch();
  return check_launch("gemm_simt_kernel");
}

}  // namespace d2r
