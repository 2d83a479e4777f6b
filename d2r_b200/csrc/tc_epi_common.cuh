// Generic epilogue helpers of the tcgen05 kernels (included inside namespace d2r::<anon> by gemm_tc.cu and
// attn_fused.cu): fast math, the per-warp staging tile and the TMA bulk tensor stores that drain it.
#pragma once

struct NoRes {};

__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// barrier of the two epilogue warps that share TMEM lane quarter q (named barriers 1..4; 0 is __syncthreads)
__device__ __forceinline__ void pair_barrier(int q) {
  asm volatile("bar.sync %0, 64;" ::"r"(q + 1) : "memory");
}

__device__ __forceinline__ float fast_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <typename T>
__device__ __forceinline__ void st_group(T* ptr, const float (&v)[8], int nvalid) {
  if (nvalid == 8 && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0) {
    store8(ptr, v);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nvalid) Elem<T>::st(ptr + i, v[i]);
  }
}

template <typename T>
__device__ __forceinline__ void ld_group(const T* ptr, float (&v)[8], int nvalid) {
  if (nvalid == 8 && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0) {
    load8(ptr, v);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (i < nvalid) ? Elem<T>::ld(ptr + i) : 0.f;
  }
}

// Register image of one row's 32 residual values, fetched ahead of the TMEM wait (fast path only).
template <typename RT> struct ResRegs {};
template <> struct ResRegs<__nv_bfloat16> { uint4 v[4]; };
template <> struct ResRegs<float> { float4 v[8]; };

template <typename RT>
__device__ __forceinline__ void prefetch_res(ResRegs<RT>& pre, const RT* rrow, int col0) {
  if constexpr (!std::is_same<RT, NoRes>::value) {
    constexpr int NV = 32 / (16 / sizeof(RT));
    using Vec = typename std::remove_reference<decltype(pre.v[0])>::type;
    const Vec* ptr = reinterpret_cast<const Vec*>(rrow + col0);
#pragma unroll
    for (int i = 0; i < NV; ++i) pre.v[i] = ptr[i];
  }
}

template <typename RT>
__device__ __forceinline__ void unpack_group(const ResRegs<RT>& pre, int g, float (&res)[8]) {
  if constexpr (std::is_same<RT, __nv_bfloat16>::value) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&pre.v[g]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(h[i]);
      res[2 * i] = f.x;
      res[2 * i + 1] = f.y;
    }
  } else if constexpr (std::is_same<RT, float>::value) {
    const float4 a = pre.v[2 * g], b = pre.v[2 * g + 1];
    res[0] = a.x; res[1] = a.y; res[2] = a.z; res[3] = a.w;
    res[4] = b.x; res[5] = b.y; res[6] = b.z; res[7] = b.w;
  }
}

// ---- staging tile: 32 rows x 64 bytes, CU_TENSOR_MAP_SWIZZLE_64B (16-byte unit u of row r lives at
//      r*64 + ((u ^ ((r >> 1) & 3)) << 4)); conflict-free for one-row-per-lane writes.
__device__ __forceinline__ void stage_unit(uint8_t* stage, int lane, int unit, uint4 val) {
  *reinterpret_cast<uint4*>(stage + lane * 64 + ((unit ^ ((lane >> 1) & 3)) << 4)) = val;
}

__device__ __forceinline__ uint4 pack8_bf16(const float (&v)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  return u;
}

__device__ __forceinline__ void tma_store_4d(const void* tmap, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// Same, but the tile is ADDED to global memory by the TMA unit (fp32): split-K / gradient accumulation
// without a single per-thread atomic instruction.
__device__ __forceinline__ void tma_reduce_add_4d(const void* tmap, const void* smem_src, int c0, int c1, int c2,
                                                  int c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Write 32 fp32 values of this lane's row (columns col0 .. col0+31) through the staging tile.
// bf16: one 64-byte row -> one store; fp32: two rounds of 16 columns.
template <typename CT, bool REDUCE = false>
__device__ __forceinline__ void tma_store_row32(const CUtensorMap* tm, uint8_t* stage, int lane, const float (&v)[32],
                                                int col0, int row0, int zi, int zo) {
  constexpr int ROUNDS = sizeof(CT) == 2 ? 1 : 2;
#pragma unroll
  for (int rd = 0; rd < ROUNDS; ++rd) {
    if (lane == 0) bulk_wait_read0();      // the previous store has finished reading the staging tile
    __syncwarp();
    if constexpr (sizeof(CT) == 2) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float t[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = v[g * 8 + i];
        stage_unit(stage, lane, g, pack8_bf16(t));
      }
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint4 val;
        val.x = __float_as_uint(v[rd * 16 + u * 4 + 0]);
        val.y = __float_as_uint(v[rd * 16 + u * 4 + 1]);
        val.z = __float_as_uint(v[rd * 16 + u * 4 + 2]);
        val.w = __float_as_uint(v[rd * 16 + u * 4 + 3]);
        stage_unit(stage, lane, u, val);
      }
    }
    fence_proxy_async();                   // generic-proxy smem writes -> visible to the async (TMA) proxy
    __syncwarp();
    if (lane == 0) {
      if constexpr (REDUCE) tma_reduce_add_4d(tm, stage, col0 + rd * 16, row0, zi, zo);
      else tma_store_4d(tm, stage, col0 + rd * 16, row0, zi, zo);
      bulk_commit();
    }
  }
}

