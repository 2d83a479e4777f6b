// Epilogue of the tcgen05 GEMM (included by gemm_tc.cu inside namespace d2r::<anon>).
//
// 8 epilogue warps.  Warp pair (q, half) shares TMEM lane quarter q (rows q*32 .. q*32+31 of the tile);
// half 0 takes the first half of the tile's 32-column chunks, half 1 the second, so two warps per
// scheduler hide each other's TMEM / shared-memory latency.
//
// Output path: accumulator chunk (tcgen05.ld, one row per thread) -> bias/act/residual in registers ->
// 64-byte-swizzled staging tile in shared memory -> ONE TMA bulk tensor store per chunk
// (cp.async.bulk.tensor ... global.shared::cta).  Measured on B200: the scattered per-thread 16-byte
// global stores of a direct epilogue cost ~30% of a K=768 GEMM; the bulk store removes them from the LSU.
// TMA clips rows >= m and columns >= n, so ragged edge tiles take the same path.  C pointers / strides
// that violate TMA's 16-byte rules, and atomic accumulation (split-K), use direct stores.
#pragma once

#include "tc_epi_common.cuh"

// One chunk: 32 accumulator columns of this lane's row -> v[32] (and d[32] for the squared difference).
//   MODE 0: v = act(alpha*acc + bias) + residual     MODE 1: d = residual - (alpha*acc + bias), v = d*d
//   MODE 2: v = alpha*acc + bias (atomic accumulation, stored by the caller)
template <typename CT, typename RT, int MODE, bool FAST>
__device__ __forceinline__ void epilogue_math(const TcParams& p, const uint32_t (&r)[32], const float* sbias,
                                              const ResRegs<RT>& pre, const RT* rrow, int col0, float (&v)[32],
                                              float (&d)[32]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int col = col0 + g * 8;
    float t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = p.alpha * __uint_as_float(r[g * 8 + i]);
    if (sbias) {
      const float4 b0 = *reinterpret_cast<const float4*>(sbias + g * 8);
      const float4 b1 = *reinterpret_cast<const float4*>(sbias + g * 8 + 4);
      t[0] += b0.x; t[1] += b0.y; t[2] += b0.z; t[3] += b0.w;
      t[4] += b1.x; t[5] += b1.y; t[6] += b1.z; t[7] += b1.w;
    }
    float res[8];
    if constexpr (!std::is_same<RT, NoRes>::value) {
      if constexpr (FAST) {
        unpack_group(pre, g, res);
      } else {
        const int nvalid = max(0, min(8, p.n - col));
        ld_group(rrow + col, res, nvalid);
      }
    }
    if constexpr (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        d[g * 8 + i] = res[i] - t[i];
        t[i] = d[g * 8 + i] * d[g * 8 + i];
      }
    } else if constexpr (MODE == 0) {
      if (p.act != D2R_ACT_NONE && (p.act_cols == 0 || col < p.act_cols)) {
        if (p.act == D2R_ACT_RELU) {
#pragma unroll
          for (int i = 0; i < 8; ++i) t[i] = fmaxf(t[i], 0.f);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) t[i] = sizeof(CT) == 2 ? fast_tanh(t[i]) : tanhf(t[i]);
        }
      }
      if constexpr (!std::is_same<RT, NoRes>::value) {
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] += res[i];
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) v[g * 8 + i] = t[i];
  }
}

// direct (non-TMA) stores of one chunk
template <typename CT, int MODE>
__device__ __forceinline__ void direct_store(const TcParams& p, CT* crow, CT* c2row, int col0, const float (&v)[32],
                                             const float (&d)[32]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int col = col0 + g * 8;
    if (col >= p.n) break;
    const int nvalid = min(8, p.n - col);
    float t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = v[g * 8 + i];
    if constexpr (MODE == 2) {
      float* dst = reinterpret_cast<float*>(crow) + col;
      if (nvalid == 8 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        // red.global.add.v4.f32: one L2 reduction per 16 bytes instead of one per element
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(t[0]), "f"(t[1]), "f"(t[2]),
                     "f"(t[3])
                     : "memory");
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(t[4]), "f"(t[5]), "f"(t[6]),
                     "f"(t[7])
                     : "memory");
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (i < nvalid) atomicAdd(dst + i, t[i]);
      }
    } else {
      if constexpr (MODE == 1) {
        float u[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) u[i] = d[g * 8 + i];
        st_group(c2row + col, u, nvalid);
      }
      st_group(crow + col, t, nvalid);
    }
  }
}

// TileWalk: which tiles this CTA visits (persistent stride) and, for a CTA pair (cta_group::2), which 128-row half
// of the 256-row pair tile it owns and where the accumulator-free signal goes (the leader CTA's barrier).
struct TileWalk {
  long long first, step;
  int m_off;        // row offset of this CTA inside the tile (0, or 128 for the odd CTA of a pair)
  int remote_cta;   // < 0: arrive on the local tmem_empty barrier; >= 0: arrive on that CTA's barrier (mapa)
};

template <int BN, typename CT, typename RT, int MODE>
__device__ __forceinline__ void epilogue_loop(const TcParams& p, const CUtensorMap* tmC, const CUtensorMap* tmC2,
                                              uint32_t tmem_base, uint64_t* tmem_full, uint64_t* tmem_empty,
                                              float* sbias_warp, uint8_t* stage, int q, int half, int lane,
                                              const TileWalk w) {
  int acc = 0;
  uint32_t acc_phase = 0;
  const bool use_tma = p.tma_store != 0;
  const bool prof = p.prof != nullptr && q == 0 && half == 0 && lane == 0;
  long long w_full = 0;
  const long long t_loop = prof ? clock64() : 0;
  for (long long t = w.first; t < p.num_tiles; t += w.step) {
    const TileCoord tc = decode_tile(p, t, BN);
    // stage this tile's bias slice in shared memory while the accumulator is still being produced
    if (p.bias) {
      const float* bias = p.bias + static_cast<long long>(tc.z) * p.bias_sz;
      for (int i = lane; i < BN; i += 32) sbias_warp[i] = (tc.n0 + i < p.n) ? __ldg(bias + tc.n0 + i) : 0.f;
    }
    __syncwarp();
    mbar_wait_timed(&tmem_full[acc], acc_phase, prof, w_full);
    tc_fence_after();
    const int row0 = tc.m0 + w.m_off + q * 32;
    const long long row = row0 + lane;
    const bool row_ok = row < p.m;
    const long long c_off = static_cast<long long>(tc.zo) * p.c_so + static_cast<long long>(tc.zi) * p.c_si +
                            row * p.ldc;
    CT* crow = reinterpret_cast<CT*>(p.c) + c_off;
    CT* c2row = MODE == 1 ? reinterpret_cast<CT*>(p.c2) + c_off : nullptr;
    const RT* rrow = nullptr;
    if constexpr (!std::is_same<RT, NoRes>::value)
      rrow = reinterpret_cast<const RT*>(p.residual) + static_cast<long long>(tc.zo) * p.r_so +
             static_cast<long long>(tc.zi) * p.r_si + row * p.ldr;
    const int ncols = min(BN, p.n - tc.n0);
    const int nchunks = (ncols + 31) >> 5;
    const int cmid = (nchunks + 1) >> 1;
    const int cb = half == 0 ? 0 : cmid;
    const int ce = half == 0 ? cmid : nchunks;
    const uint32_t taddr = tmem_base + static_cast<uint32_t>(acc * BN) + (static_cast<uint32_t>(q * 32) << 16);
    const float* sb = p.bias ? sbias_warp : nullptr;
    // residual fast path: whole tile inside the matrix and every row pointer 16-byte aligned (warp-uniform)
    const bool r_ok = (reinterpret_cast<uintptr_t>(rrow) & 15) == 0;
    const bool fast = (tc.n0 + BN <= p.n) && __all_sync(0xffffffffu, r_ok && row_ok);
    if constexpr (MODE == 3 || MODE == 4) {
      // ---- row softmax fused into the score GEMM (MODE 3) / its backward fused into the dP GEMM (MODE 4).
      // The whole key dimension is one N tile (n <= BN), so the two warps that share a TMEM lane quarter hold a
      // complete row between them: S / dP never touch HBM.
      const float sc = p.alpha * 1.4426950408889634f;       // exp(alpha x) = 2^(alpha log2(e) x)
      if (nchunks <= 4) {
        // Each warp keeps ITS chunks (<= 2) in registers, reduces them, and swaps one partial result per row with
        // its partner warp through shared memory (two slots, alternating per tile: the pair barrier of tile i+1
        // cannot complete before the partner has read the slot of tile i).
        float* xch = sbias_warp + (acc & 1) * 64;
        const float* xch_peer = xch + (half == 0 ? 4 : -4) * (BN < 128 ? 128 : BN);
        if constexpr (MODE == 3) {
          float e[64];
          float ml = -INFINITY;
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int c = cb + cc;
            if (c < ce) {
              uint32_t ra[32];
              tmem_ld32(taddr + c * 32, ra);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float x = (c * 32 + i < ncols) ? sc * __uint_as_float(ra[i]) : -INFINITY;
                e[cc * 32 + i] = x;
                ml = fmaxf(ml, x);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) e[cc * 32 + i] = -INFINITY;
            }
          }
          const float ms = ml == -INFINITY ? 0.f : ml;       // a warp without a valid column contributes 0
          float sl = 0.f;
#pragma unroll
          for (int i = 0; i < 64; ++i) {
            e[i] = ex2_ftz(e[i] - ms);
            sl += e[i];
          }
          *reinterpret_cast<float2*>(xch + lane * 2) = make_float2(ml, sl);
          pair_barrier(q);
          const float2 pp = *reinterpret_cast<const float2*>(xch_peer + lane * 2);
          const float mx = fmaxf(ml, pp.x);
          const float mine = ex2_ftz(ml - mx);               // 2^(-inf) = 0
          const float f = mine / (sl * mine + pp.y * ex2_ftz(pp.x - mx));
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int c = cb + cc;
            if (c < ce) {
              float v[32], d[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = e[cc * 32 + i] * f;
              if (use_tma) tma_store_row32<CT>(tmC, stage, lane, v, c * 32, row0, tc.zi, tc.zo);
              else if (row_ok) direct_store<CT, 0>(p, crow, c2row, c * 32, v, d);
            }
          }
        } else {
          // P of this warp's <= 2 chunks is fetched once and kept packed (bf16, 32 registers) across the exchange
          float part = 0.f;
          uint4 pk[2][4];
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {           // pass 1: this warp's share of sum_j dP_j P_j
            const int c = cb + cc;
            if (c < ce) {
              uint32_t ra[32];
              tmem_ld32(taddr + c * 32, ra);
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                float pr[8];
                if (row_ok) {
                  ld_group(rrow + c * 32 + g * 8, pr, max(0, min(8, ncols - c * 32 - g * 8)));
                } else {
#pragma unroll
                  for (int i = 0; i < 8; ++i) pr[i] = 0.f;
                }
                pk[cc][g] = pack8_bf16(pr);          // exact: P is bf16 in memory
              }
              tmem_ld_wait();
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&pk[cc][g]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 f = __bfloat1622float2(h[i]);
                  part = fmaf(p.alpha * __uint_as_float(ra[g * 8 + 2 * i]), f.x, part);
                  part = fmaf(p.alpha * __uint_as_float(ra[g * 8 + 2 * i + 1]), f.y, part);
                }
              }
            }
          }
          xch[lane] = part;
          pair_barrier(q);
          const float tot = part + xch_peer[lane];
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {           // pass 2: dS = P * (alpha dP - total)
            const int c = cb + cc;
            if (c < ce) {
              uint32_t ra[32];
              float v[32], d[32];
              tmem_ld32(taddr + c * 32, ra);
              tmem_ld_wait();
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&pk[cc][g]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 f = __bfloat1622float2(h[i]);
                  v[g * 8 + 2 * i] = f.x * (p.alpha * __uint_as_float(ra[g * 8 + 2 * i]) - tot);
                  v[g * 8 + 2 * i + 1] = f.y * (p.alpha * __uint_as_float(ra[g * 8 + 2 * i + 1]) - tot);
                }
              }
              if (use_tma) tma_store_row32<CT>(tmC, stage, lane, v, c * 32, row0, tc.zi, tc.zo);
              else if (row_ok) direct_store<CT, 0>(p, crow, c2row, c * 32, v, d);
            }
          }
        }
      } else {
      // wide rows (128 < n <= 256): both warps compute the statistics of the full row (TMEM re-reads), then each
      // writes its own half
      float stat0 = MODE == 3 ? -INFINITY : 0.f, stat1 = 0.f;
#pragma unroll 1
      for (int c = 0; c < nchunks; ++c) {          // pass 1: row max (fwd) / sum dP*P (bwd)
        uint32_t ra[32];
        tmem_ld32(taddr + c * 32, ra);
        tmem_ld_wait();
        if constexpr (MODE == 3) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i < ncols) stat0 = fmaxf(stat0, sc * __uint_as_float(ra[i]));
        } else if (row_ok) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float pr[8];
            ld_group(rrow + c * 32 + g * 8, pr, max(0, min(8, ncols - c * 32 - g * 8)));
#pragma unroll
            for (int i = 0; i < 8; ++i) stat0 = fmaf(p.alpha * __uint_as_float(ra[g * 8 + i]), pr[i], stat0);
          }
        }
      }
      if constexpr (MODE == 3) {
#pragma unroll 1
        for (int c = 0; c < nchunks; ++c) {        // pass 2: sum of exponentials
          uint32_t ra[32];
          tmem_ld32(taddr + c * 32, ra);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i < ncols) stat1 += ex2_ftz(sc * __uint_as_float(ra[i]) - stat0);
        }
        stat1 = 1.f / stat1;
      }
#pragma unroll 1
      for (int c = cb; c < ce; ++c) {              // pass 3: this warp's half of the row -> P (or dS)
        uint32_t ra[32];
        float v[32], d[32];
        tmem_ld32(taddr + c * 32, ra);
        tmem_ld_wait();
        if constexpr (MODE == 3) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            v[i] = (c * 32 + i < ncols) ? ex2_ftz(sc * __uint_as_float(ra[i]) - stat0) * stat1 : 0.f;
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float pr[8];
            if (row_ok) ld_group(rrow + c * 32 + g * 8, pr, max(0, min(8, ncols - c * 32 - g * 8)));
#pragma unroll
            for (int i = 0; i < 8; ++i)
              v[g * 8 + i] = row_ok ? pr[i] * (p.alpha * __uint_as_float(ra[g * 8 + i]) - stat0) : 0.f;
          }
        }
        if (use_tma) tma_store_row32<CT>(tmC, stage, lane, v, c * 32, row0, tc.zi, tc.zo);
        else if (row_ok) direct_store<CT, 0>(p, crow, c2row, c * 32, v, d);
      }
      }
    } else {
#pragma unroll 1
    for (int c = cb; c < ce; ++c) {
      if (p.debug & 2) continue;
      uint32_t ra[32];
      ResRegs<RT> pa;
      float v[32], d[32];
      const int col0 = tc.n0 + c * 32;
      tmem_ld32(taddr + c * 32, ra);
      if (fast) prefetch_res<RT>(pa, rrow, col0);
      tmem_ld_wait();
      if (fast) {
        epilogue_math<CT, RT, MODE, true>(p, ra, sb ? sb + c * 32 : nullptr, pa, rrow, col0, v, d);
      } else if (row_ok) {
        epilogue_math<CT, RT, MODE, false>(p, ra, sb ? sb + c * 32 : nullptr, pa, rrow, col0, v, d);
      }
      if (p.debug & 1) continue;
      if (use_tma) {
        if constexpr (MODE == 1) tma_store_row32<CT>(tmC2, stage, lane, d, col0, row0, tc.zi, tc.zo);
        if constexpr (MODE == 2) {
          if constexpr (sizeof(CT) == 4) tma_store_row32<CT, true>(tmC, stage, lane, v, col0, row0, tc.zi, tc.zo);
        } else {
          tma_store_row32<CT>(tmC, stage, lane, v, col0, row0, tc.zi, tc.zo);
        }
      } else if (row_ok) {
        direct_store<CT, MODE>(p, crow, c2row, col0, v, d);
      }
    }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (w.remote_cta < 0) mbar_arrive(&tmem_empty[acc]);
      else mbar_arrive_cluster(&tmem_empty[acc], static_cast<uint32_t>(w.remote_cta));
    }
    if (++acc == num_acc(BN)) {
      acc = 0;
      acc_phase ^= 1;
    }
  }
  if (prof) {
    long long* rec = p.prof + static_cast<long long>(blockIdx.x) * kProfSlots;
    rec[PS_EPI_WAIT_FULL] = w_full;
    rec[PS_EPI_LOOP] = clock64() - t_loop;
  }
  if (lane == 0) bulk_wait_all();          // all bulk stores of this warp are complete before the CTA exits
  __syncwarp();
}

enum EpiVariant {
  EV_F32 = 0, EV_F32_RF32, EV_F32_RBF16, EV_BF16, EV_BF16_RF32, EV_BF16_RBF16, EV_SQ_F32, EV_SQ_BF16, EV_ATOMIC,
  EV_SOFTMAX_BF16, EV_SOFTMAX_BWD_BF16
};
