// Fused attention forward for the cells of the routed stack: ONE kernel computes
//
//     P = softmax_row(alpha * Q K^T)        (kept in bf16, also written out for the backward)
//     O = P V  (+ residual)                 or, for the alignment cell,  d = residual - P V,  O2 = d,  O = d * d
//
// for every (sample, head, 128-row query tile) "unit".  The score matrix lives in TMEM only and the probabilities
// go from registers straight into the 128B-swizzled shared-memory tile that the second tcgen05.mma reads as its A
// operand -- the composed path (models/SelfAttention.py:33-39, models/XModules.py:300-310, models/Cells.py:244-246
// as two d2r_gemm launches) wrote P to HBM and read it back, and every launch drained and refilled the GPU.
//
// Shapes served (bf16 only): keys Lc <= 128 (one N tile: a whole score row is in one accumulator), any Lq
// (tiles of 128 rows), head dim 48 (16-head self-attention, SelfAttention.py:27-42) or a multiple of 64 such as 768
// (the single-head cross-modal attentions).  Operands are addressed like d2r_gemm's head-strided batches:
// X[b, row, h * hd + col] with a row stride and (rows * row stride) per sample.
//
// Warp roles (320 threads, one CTA per SM, persistent over units u = blockIdx.x, += gridDim.x):
//   warp 0      TMA producer: per unit the Q / K k-blocks of phase 1, then the V tiles of phase 2, through one ring
//   warp 1      MMA issuer:   S(u+1) is issued BEFORE P V(u), so the softmax of unit u overlaps the next unit's
//                             score product (two S accumulators, two P tiles, two O accumulators: 512 TMEM columns)
//   warps 2..5  softmax (one thread per score row: TMEM -> registers -> bf16 P tile in shared memory + TMA store of P)
//   warps 6..9  output epilogue (P V accumulator tiles -> (+ residual | squared difference) -> TMA store)
// Barriers: full/empty[stage] (TMA <-> MMA), s_full/s_free[2] (MMA <-> softmax), p_ready[2] (softmax -> MMA),
//           o_full/o_empty[2] (MMA <-> output epilogue).
#include <cuda.h>

#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"

namespace d2r {
namespace {

#include "tc_host.cuh"
#include "tc_epi_common.cuh"

constexpr int kQTile = 128 * 64 * 2;    // one [128 x 64] bf16 K-major operand block = 16 KB
constexpr int kAtom = 64 * 64 * 2;      // one [64 x 64] bf16 swizzle atom = 8 KB
constexpr uint32_t kColS = 0, kColO = 256;   // TMEM: S0 | S1 | O0 | O1, 128 fp32 columns each

template <int BNS>
struct AfCfg {
  static constexpr int STAGE_BYTES = kQTile + BNS * 128;            // Q k-block + K k-block (24 / 32 KB)
  static constexpr int STAGES = BNS == 64 ? 6 : 4;
  static constexpr int P_BYTES = (BNS / 64) * kQTile;               // one P tile [128 x BNS] bf16
  static constexpr int STORE_BYTES = 8 * 4096;                      // one [32 x 128 B] staging tile per epilogue warp
  static constexpr int BAR_BYTES = 512;
  static constexpr int XCH_BYTES = 8 * 64 * 4;                      // softmax statistics exchanged inside a warp pair
  // no 1 KB alignment slack: the kernel declares its dynamic shared memory __align__(1024) (and traps if the base is
  // not aligned); the 128-key configuration needs 226.5 KB of the 227 KB a CTA may have
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * P_BYTES + STORE_BYTES + BAR_BYTES + XCH_BYTES;
  static_assert(SMEM_BYTES <= 227 * 1024, "attn_fwd: shared memory budget");
};

struct AfParams {
  int Lq, Lc, hd, heads, q_tiles, units;
  int nkb;         // k-blocks of the score product (ceil(hd / 64))
  int ksteps2;     // UMMA K steps (16 keys each) of the P V product that hold valid keys
  int nt, n_tiles; // output tile width (64 | 128 columns of the head dim) and count
  float sc;        // alpha * log2(e)
  int mode;        // 0: O = P V (+ residual);  1: d = residual - P V, O2 = d, O = d * d
  const __nv_bfloat16* residual;
  long long r_ld, r_sb, r_sh;   // residual: row stride, sample stride, head stride (elements)
};

struct Unit {
  int b, h, qt;
};
__device__ __forceinline__ Unit decode_unit(const AfParams& p, int u) {
  Unit x;
  x.qt = u % p.q_tiles;
  const int bh = u / p.q_tiles;
  x.h = bh % p.heads;
  x.b = bh / p.heads;
  return x;
}

// SPLIT (small head dims, one output tile per unit): warps 2..5 do the softmax, warps 6..9 the output tiles, so the
// softmax of unit i+1 overlaps the output of unit i.  !SPLIT (head dim 768: six output tiles per unit, the output
// dominates): all eight warps do both, a warp pair sharing a lane quarter splits the columns.
template <int BNS, bool SPLIT>
__global__ void __launch_bounds__(kTcThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmP,
                const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmO2, const AfParams p) {
  using Cfg = AfCfg<BNS>;
  constexpr int KCB = BNS / 64;        // 64-key blocks of the P V product
  constexpr int NCH = BNS / 32;        // 32-column chunks of a score row
  constexpr int CPW = NCH / 2;         // chunks per warp of a pair (!SPLIT)
  constexpr int kArrive = SPLIT ? 4 : 8;   // warps that consume a score / output accumulator
  extern __shared__ __align__(1024) uint8_t smem_fwd[];
  if (smem_u32(smem_fwd) & 1023u) __trap();   // 128B-swizzled tiles must start on a 1024-byte boundary
  uint8_t* smem = smem_fwd;
  uint8_t* pbuf = smem + Cfg::STAGES * Cfg::STAGE_BYTES;            // 2 x P_BYTES, 1024-aligned
  uint8_t* store_stage = pbuf + 2 * Cfg::P_BYTES;                   // [8][4096], 1024-aligned
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(store_stage + Cfg::STORE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* s_full = empty_bar + Cfg::STAGES;
  uint64_t* s_free = s_full + 2;
  uint64_t* p_ready = s_free + 2;
  uint64_t* o_full = p_ready + 2;
  uint64_t* o_empty = o_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 2);
  float* xch = reinterpret_cast<float*>(store_stage + Cfg::STORE_BYTES + Cfg::BAR_BYTES);   // [8][64]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmP);
    tma_prefetch_desc(&tmO);
    if (p.mode == 1) tma_prefetch_desc(&tmO2);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&s_full[a], 1);
      mbar_init(&s_free[a], kArrive);
      mbar_init(&p_ready[a], kArrive);
      mbar_init(&o_full[a], 1);
      mbar_init(&o_empty[a], kArrive);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  const int n_mine = p.units > (int)blockIdx.x ? (p.units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      auto advance = [&]() {
        if (++stage == Cfg::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      };
      auto load_qk = [&](int i) {
        const Unit u = decode_unit(p, (int)blockIdx.x + i * (int)gridDim.x);
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          tma_load_4d(sa, &tmQ, &full_bar[stage], kb * 64, u.qt * 128, u.h, u.b);
          tma_load_4d(sa + kQTile, &tmK, &full_bar[stage], kb * 64, 0, u.h, u.b);
          advance();
        }
      };
      auto load_v = [&](int i) {
        const Unit u = decode_unit(p, (int)blockIdx.x + i * (int)gridDim.x);
        const int atoms = p.nt / 64;
        for (int t = 0; t < p.n_tiles; ++t) {
          for (int kc = 0; kc < KCB; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], atoms * kAtom);
            uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
            for (int a = 0; a < atoms; ++a)
              tma_load_4d(sa + a * kAtom, &tmV, &full_bar[stage], t * p.nt + 64 * a, kc * 64, u.h, u.b);
            advance();
          }
        }
      };
      if (n_mine > 0) load_qk(0);
      for (int i = 0; i < n_mine; ++i) {
        if (i + 1 < n_mine) load_qk(i + 1);
        load_v(i);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, BNS, 0, 0);
      const uint32_t idesc_o = make_idesc_bf16(128, p.nt, 0, 1);      // A = P (K-major), B = V (MN-major)
      int stage = 0;
      uint32_t phase = 0;
      int oc = 0;                                                      // output tiles issued so far
      auto advance = [&]() {
        if (++stage == Cfg::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      };
      auto issue_s = [&](int i) {
        const int sb = i & 1;
        mbar_wait(&s_free[sb], (static_cast<uint32_t>(i >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + kColS + static_cast<uint32_t>(sb * 128);
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sk = sa + kQTile;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(d_tmem, make_smem_desc_sw128(sa + k * 32, 16, 1024), make_smem_desc_sw128(sk + k * 32, 16, 1024),
                      idesc_s, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          advance();
        }
        umma_commit(&s_full[sb]);
      };
      auto issue_pv = [&](int i) {
        const int sb = i & 1;
        mbar_wait(&p_ready[sb], static_cast<uint32_t>(i >> 1) & 1u);
        tc_fence_after();
        const uint32_t pa = smem_u32(pbuf + sb * Cfg::P_BYTES);
        for (int t = 0; t < p.n_tiles; ++t) {
          const int ob = oc & 1;
          mbar_wait(&o_empty[ob], (static_cast<uint32_t>(oc >> 1) & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + kColO + static_cast<uint32_t>(ob * 128);
          for (int kc = 0; kc < KCB; ++kc) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sv = smem_u32(smem + stage * Cfg::STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (kc * 4 + k < p.ksteps2)
                umma_bf16(d_tmem, make_smem_desc_sw128(pa + kc * kQTile + k * 32, 16, 1024),
                          make_smem_desc_sw128(sv + k * 2048, kAtom, 1024), idesc_o, (kc > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&empty_bar[stage]);
            advance();
          }
          umma_commit(&o_full[ob]);
          ++oc;
        }
      };
      if (n_mine > 0) issue_s(0);
      for (int i = 0; i < n_mine; ++i) {
        if (i + 1 < n_mine) issue_s(i + 1);
        issue_pv(i);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------- softmax + output epilogue (warps 2..9)
    const int q = warp & 3;                     // TMEM lane quarter of this warp
    const int ew = warp - 2;
    const int half = ew >> 2;                   // SPLIT: 0 = softmax warp, 1 = output warp; else column half of the pair
    const int r = q * 32 + lane;                // row of the unit's 128-row tile
    uint8_t* stg = store_stage + ew * 4096;
    int oc = 0;

    // output tiles of one unit: this warp drains 64-column chunk pairs [pair0, pair0 + npairs) of every tile
    auto out_tiles = [&](const Unit& u, int pair0, int npairs) {
      const int row0 = u.qt * 128 + q * 32;
      const bool row_ok = row0 + lane < p.Lq;
      const __nv_bfloat16* rrow = nullptr;
      if (p.residual)
        rrow = p.residual + static_cast<long long>(u.b) * p.r_sb + static_cast<long long>(u.h) * p.r_sh +
               static_cast<long long>(row0 + lane) * p.r_ld;
      for (int t = 0; t < p.n_tiles; ++t) {
        const int ob = oc & 1;
        for (int pi = 0; pi < npairs; ++pi) {
          const int c0 = (pair0 + pi) * 2;
          // The residual of the pair is fetched BEFORE the accumulator wait / while the previous pair drains: eight
          // independent 16-byte loads per thread in flight.  (First version: one load at a time behind the wait,
          // each followed by its unpack -- ncu showed 44 % of the stall samples on those LDG.128.)
          uint4 rr[2][4];
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int col0 = t * p.nt + (c0 + cc) * 32;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              // hd is a multiple of 8: a group of 8 columns is entirely inside or outside the head
              const bool ok = rrow != nullptr && row_ok && col0 + g * 8 < p.hd;
              rr[cc][g] = ok ? __ldg(reinterpret_cast<const uint4*>(rrow + col0 + g * 8)) : make_uint4(0u, 0u, 0u, 0u);
            }
          }
          if (pi == 0) {
            mbar_wait(&o_full[ob], static_cast<uint32_t>(oc >> 1) & 1u);
            tc_fence_after();
          }
          const int colp = t * p.nt + c0 * 32;      // first column of the pair inside the head
          if (colp >= p.hd) continue;
          // one pass per output tensor (the squared-difference epilogue writes d and d*d): chunk by chunk from TMEM
          // into the 128B-swizzled staging tile, then ONE 64-column TMA store
          const int npass = p.mode == 1 ? 2 : 1;
          for (int pass = 0; pass < npass; ++pass) {
            if (lane == 0) bulk_wait_read0();       // the previous store has finished reading the staging tile
            __syncwarp();
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
              uint32_t ra[32];
              tmem_ld32(tmem_base + kColO + static_cast<uint32_t>(ob * 128) + (static_cast<uint32_t>(q * 32) << 16) +
                            (c0 + cc) * 32, ra);
              tmem_ld_wait();
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&rr[cc][g]);
                float t8[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 fl = __bfloat1622float2(hh[j]);          // (zero without residual)
                  const float a0 = __uint_as_float(ra[g * 8 + 2 * j]), a1 = __uint_as_float(ra[g * 8 + 2 * j + 1]);
                  if (p.mode == 1) {
                    const float d0 = fl.x - a0, d1 = fl.y - a1;
                    t8[2 * j] = pass == 0 ? d0 : d0 * d0;
                    t8[2 * j + 1] = pass == 0 ? d1 : d1 * d1;
                  } else {
                    t8[2 * j] = a0 + fl.x;
                    t8[2 * j + 1] = a1 + fl.y;
                  }
                }
                *reinterpret_cast<uint4*>(stg + lane * 128 + (((cc * 4 + g) ^ (lane & 7)) << 4)) = pack8_bf16(t8);
              }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_4d((p.mode == 1 && pass == 0) ? &tmO2 : &tmO, stg, colp, row0, u.h, u.b);
              bulk_commit();
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_empty[ob]);
        ++oc;
      }
    };

    if constexpr (SPLIT) {
      if (half == 0) {
        // ---- softmax warps: one thread per score row
        for (int i = 0; i < n_mine; ++i) {
          const Unit u = decode_unit(p, (int)blockIdx.x + i * (int)gridDim.x);
          const int sb = i & 1;
          uint8_t* pb = pbuf + sb * Cfg::P_BYTES;
          mbar_wait(&s_full[sb], static_cast<uint32_t>(i >> 1) & 1u);
          tc_fence_after();
          const uint32_t ts = tmem_base + kColS + static_cast<uint32_t>(sb * 128) + (static_cast<uint32_t>(q * 32) << 16);
          float e[NCH * 32];
          float ml = -INFINITY;
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            uint32_t ra[32];
            tmem_ld32(ts + c * 32, ra);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float x = (c * 32 + j < p.Lc) ? p.sc * __uint_as_float(ra[j]) : -INFINITY;
              e[c * 32 + j] = x;
              ml = fmaxf(ml, x);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_free[sb]);  // the score accumulator may be overwritten (unit i + 2)
          float sl = 0.f;
#pragma unroll
          for (int j = 0; j < NCH * 32; ++j) {
            e[j] = ex2_ftz(e[j] - ml);              // Lc >= 1: the row maximum is finite
            sl += e[j];
          }
          const float f = 1.f / sl;
          // the bulk stores that read this warp's rows of this P tile two units ago were issued by this lane
          if (lane == 0) bulk_wait_read0();
          __syncwarp();
#pragma unroll
          for (int g = 0; g < NCH * 4; ++g) {
            float t8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) t8[j] = e[g * 8 + j] * f;
            const int col0 = g * 8;
            *reinterpret_cast<uint4*>(pb + (col0 >> 6) * kQTile + r * 128 + ((((col0 & 63) >> 3) ^ (r & 7)) << 4)) =
                pack8_bf16(t8);
          }
          fence_proxy_async();                      // generic-proxy writes of P -> visible to tcgen05.mma / TMA
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&p_ready[sb]);
#pragma unroll
            for (int a = 0; a < KCB; ++a)
              tma_store_4d(&tmP, pb + a * kQTile + q * 32 * 128, a * 64, u.qt * 128 + q * 32, u.h, u.b);
            bulk_commit();
          }
        }
      } else {
        // ---- output warps
        for (int i = 0; i < n_mine; ++i) {
          const Unit u = decode_unit(p, (int)blockIdx.x + i * (int)gridDim.x);
          out_tiles(u, 0, p.nt >> 6);
        }
      }
    } else {
      float* xch_mine = xch + ew * 64;
      const float* xch_peer = xch + (half == 0 ? ew + 4 : ew - 4) * 64;
      for (int i = 0; i < n_mine; ++i) {
        const Unit u = decode_unit(p, (int)blockIdx.x + i * (int)gridDim.x);
        const int sb = i & 1;
        uint8_t* pb = pbuf + sb * Cfg::P_BYTES;
        // ---- softmax of this warp pair's 32 rows; this warp owns CPW chunks of 32 columns
        mbar_wait(&s_full[sb], static_cast<uint32_t>(i >> 1) & 1u);
        tc_fence_after();
        const uint32_t ts = tmem_base + kColS + static_cast<uint32_t>(sb * 128) + (static_cast<uint32_t>(q * 32) << 16);
        float e[CPW * 32];
        float ml = -INFINITY;
#pragma unroll
        for (int cc = 0; cc < CPW; ++cc) {
          const int c = half * CPW + cc;
          uint32_t ra[32];
          tmem_ld32(ts + c * 32, ra);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float x = (c * 32 + j < p.Lc) ? p.sc * __uint_as_float(ra[j]) : -INFINITY;
            e[cc * 32 + j] = x;
            ml = fmaxf(ml, x);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[sb]);  // the score accumulator may be overwritten (unit i + 2)
        const float ms = ml == -INFINITY ? 0.f : ml;
        float sl = 0.f;
#pragma unroll
        for (int j = 0; j < CPW * 32; ++j) {
          e[j] = ex2_ftz(e[j] - ms);
          sl += e[j];
        }
        *reinterpret_cast<float2*>(xch_mine + lane * 2) = make_float2(ml, sl);
        // the bulk stores that read this P tile two units ago were issued by this lane: drained before anyone rewrites
        if (half == 0 && lane == 0) bulk_wait_read0();
        pair_barrier(q);
        const float2 pp = *reinterpret_cast<const float2*>(xch_peer + lane * 2);
        const float mx = fmaxf(ml, pp.x);
        const float mine = ex2_ftz(ml - mx);
        const float f = mine / (sl * mine + pp.y * ex2_ftz(pp.x - mx));
#pragma unroll
        for (int cc = 0; cc < CPW; ++cc) {
          const int c = half * CPW + cc;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float t8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) t8[j] = e[cc * 32 + g * 8 + j] * f;
            const int col0 = c * 32 + g * 8;
            const int unit16 = (col0 & 63) >> 3;
            *reinterpret_cast<uint4*>(pb + (col0 >> 6) * kQTile + r * 128 + ((unit16 ^ (r & 7)) << 4)) = pack8_bf16(t8);
          }
        }
        fence_proxy_async();                      // generic-proxy writes of P -> visible to tcgen05.mma / TMA
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[sb]);
        pair_barrier(q);                          // both halves of this quarter's rows are in shared memory
        if (half == 0 && lane == 0) {
#pragma unroll
          for (int a = 0; a < KCB; ++a)
            tma_store_4d(&tmP, pb + a * kQTile + q * 32 * 128, a * 64, u.qt * 128 + q * 32, u.h, u.b);
          bulk_commit();
        }
        // ---- output tiles: each warp of the pair drains one 64-column half of every 128-column tile
        //      (64-column tiles, i.e. small heads, always take the SPLIT kernel)
        out_tiles(u, half, 1);
      }
    }
    if (lane == 0) bulk_wait_all();
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int BNS, bool SPLIT>
int launch_attn_fwd(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const CUtensorMap& tmP,
                    const CUtensorMap& tmO, const CUtensorMap& tmO2, const AfParams& p, cudaStream_t stream) {
  using Cfg = AfCfg<BNS>;
  auto kern = attn_fwd_kernel<BNS, SPLIT>;
  static std::atomic<int> attr_set[kMaxDevices];
  const int dev = current_device();
  if (!attr_set[dev].load(std::memory_order_acquire)) {
    D2R_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set[dev].store(1, std::memory_order_release);
  }
  const int grid = p.units < num_sms() ? p.units : num_sms();
  D2R_CUDA_OK(launch_pdl(kern, dim3((unsigned)grid), Cfg::SMEM_BYTES, stream, tmQ, tmK, tmV, tmP, tmO, tmO2, p));
  count_launch();
  return check_launch("attn_fwd_kernel");
}


// ===================================================================== backward
// One kernel per attention for Lq <= 128, Lc <= 128 (one tile per (sample, head)):
//
//   dP  = sign * dO V^T                      scores accumulator (TMEM), k = head dim
//   dS  = alpha * P o (dP - rowsum(dP o P))  registers -> bf16 tile in shared memory ([q, c], 128B swizzle);
//                                            P (read once from HBM) goes into a second tile
//   dV  = sign * P^T dO      A = P tile read MN-major,  B = dO streamed again as [q rows, n columns]
//   dQ  = dS K               A = dS tile read K-major,  B = K streamed
//   dK  = dS^T Q             A = dS tile read MN-major, B = Q streamed
//
// The composed path ran four d2r_gemm launches (models/SelfAttention.py:33-39 backward etc.), each re-reading
// its [B*Lq, D] operands from HBM and round-tripping dS.  Same warp roles and barrier scheme as the forward;
// the output products share two TMEM accumulators, one tile of NT head-dim columns at a time.
struct AbParams {
  int Lq, Lc, hd, heads, units;
  int nkb;            // k-blocks of the dP product (ceil(hd / 64))
  int ksteps_q;       // valid UMMA K steps when the contraction runs over queries (ceil(Lq / 16))
  int ksteps_c;       // ... over keys (ceil(Lc / 16))
  int nt, n_tiles;    // output tile width (64 | 128) and tiles per output
  float alpha, sign;
  const __nv_bfloat16* P;
  long long p_ld;
};

template <int BNS>
struct AbCfg {
  static constexpr int STAGE_BYTES = 2 * kQTile;                     // 32 KB: dO block + V block, or one B tile
  static constexpr int STAGES = 4;
  static constexpr int TILE_BYTES = 2 * kQTile;                      // [128 x 128] bf16, two 64-column atoms
  static constexpr int STORE_BYTES = 8 * 2048;
  static constexpr int BAR_BYTES = 512;
  static constexpr int XCH_BYTES = 8 * 128 * 4;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * TILE_BYTES + STORE_BYTES + BAR_BYTES + XCH_BYTES + 1024;
};

template <int BNS>
__global__ void __launch_bounds__(kTcThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmDO, const __grid_constant__ CUtensorMap tmV,
                const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmQ,
                const __grid_constant__ CUtensorMap tmDV, const __grid_constant__ CUtensorMap tmDQ,
                const __grid_constant__ CUtensorMap tmDK, const AbParams p) {
  using Cfg = AbCfg<BNS>;
  constexpr int NCH = BNS / 32;
  constexpr int CPW = NCH / 2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* ptile = smem + Cfg::STAGES * Cfg::STAGE_BYTES;           // P  [128 q x 128 c]
  uint8_t* dstile = ptile + Cfg::TILE_BYTES;                        // dS [128 q x 128 c]
  uint8_t* store_stage = dstile + Cfg::TILE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(store_stage + Cfg::STORE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* s_full = empty_bar + Cfg::STAGES;
  uint64_t* s_free = s_full + 1;
  uint64_t* ds_ready = s_free + 1;
  uint64_t* o_full = ds_ready + 1;
  uint64_t* o_empty = o_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 2);
  float* xch = reinterpret_cast<float*>(store_stage + Cfg::STORE_BYTES + Cfg::BAR_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmDV);
    tma_prefetch_desc(&tmDQ);
    tma_prefetch_desc(&tmDK);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 8);
    mbar_init(ds_ready, 8);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&o_full[a], 1);
      mbar_init(&o_empty[a], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  // the tiles' second 64-column atom is never written when there are at most 64 keys: it must read as zero
  for (int i = threadIdx.x; i < 2 * Cfg::TILE_BYTES / 16; i += kTcThreads)
    reinterpret_cast<uint4*>(ptile)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  const int n_mine = p.units > (int)blockIdx.x ? (p.units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int atoms = p.nt / 64;

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      auto advance = [&]() {
        if (++stage == Cfg::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      };
      auto load_dp = [&](int i) {
        const int u = (int)blockIdx.x + i * (int)gridDim.x;
        const int h = u % p.heads, b = u / p.heads;
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], kQTile + BNS * 128);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          tma_load_4d(sa, &tmDO, &full_bar[stage], kb * 64, 0, h, b);
          tma_load_4d(sa + kQTile, &tmV, &full_bar[stage], kb * 64, 0, h, b);
          advance();
        }
      };
      auto load_outs = [&](int i) {
        const int u = (int)blockIdx.x + i * (int)gridDim.x;
        const int h = u % p.heads, b = u / p.heads;
        for (int o = 0; o < 3; ++o) {
          const CUtensorMap* tm = o == 0 ? &tmDO : (o == 1 ? &tmK : &tmQ);
          for (int t = 0; t < p.n_tiles; ++t) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], atoms * kQTile);
            uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
            for (int a = 0; a < atoms; ++a) tma_load_4d(sa + a * kQTile, tm, &full_bar[stage], t * p.nt + 64 * a, 0, h, b);
            advance();
          }
        }
      };
      if (n_mine > 0) load_dp(0);
      for (int i = 0; i < n_mine; ++i) {
        load_outs(i);
        if (i + 1 < n_mine) load_dp(i + 1);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_dp = make_idesc_bf16(128, BNS, 0, 0);
      const uint32_t idesc_mn = make_idesc_bf16(128, p.nt, 1, 1);     // dV, dK: A tile read MN-major
      const uint32_t idesc_k = make_idesc_bf16(128, p.nt, 0, 1);      // dQ: A tile read K-major
      const uint32_t pt = smem_u32(ptile), dt = smem_u32(dstile);
      int stage = 0;
      uint32_t phase = 0;
      int oc = 0;
      auto advance = [&]() {
        if (++stage == Cfg::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      };
      auto issue_dp = [&](int i) {
        mbar_wait(s_free, (static_cast<uint32_t>(i) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + kColS;
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sv = sa + kQTile;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(d_tmem, make_smem_desc_sw128(sa + k * 32, 16, 1024), make_smem_desc_sw128(sv + k * 32, 16, 1024),
                      idesc_dp, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          advance();
        }
        umma_commit(s_full);
      };
      auto issue_outs = [&](int i) {
        mbar_wait(ds_ready, static_cast<uint32_t>(i) & 1u);
        tc_fence_after();
        for (int o = 0; o < 3; ++o) {
          const int ksteps = o == 1 ? p.ksteps_c : p.ksteps_q;
          for (int t = 0; t < p.n_tiles; ++t) {
            const int ob = oc & 1;
            mbar_wait(&o_empty[ob], (static_cast<uint32_t>(oc >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + kColO + static_cast<uint32_t>(ob * 128);
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sb = smem_u32(smem + stage * Cfg::STAGE_BYTES);
            for (int k = 0; k < ksteps; ++k) {
              // B tile: [128 contraction rows x 64-column atoms], atoms 16 KB apart, 16 rows = 2048 B per K step
              const uint64_t db = make_smem_desc_sw128(sb + k * 2048, kQTile, 1024);
              uint64_t da;
              if (o == 0) da = make_smem_desc_sw128(pt + k * 2048, kQTile, 1024);          // P^T  (M = keys)
              else if (o == 2) da = make_smem_desc_sw128(dt + k * 2048, kQTile, 1024);     // dS^T (M = keys)
              else da = make_smem_desc_sw128(dt + (k >> 2) * kQTile + (k & 3) * 32, 16, 1024);   // dS (M = queries)
              umma_bf16(d_tmem, da, db, o == 1 ? idesc_k : idesc_mn, k > 0 ? 1u : 0u);
            }
            umma_commit(&empty_bar[stage]);
            advance();
            umma_commit(&o_full[ob]);
            ++oc;
          }
        }
      };
      if (n_mine > 0) issue_dp(0);
      for (int i = 0; i < n_mine; ++i) {
        issue_outs(i);
        if (i + 1 < n_mine) issue_dp(i + 1);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------- softmax backward + output epilogue
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int ew = warp - 2;
    uint8_t* stg = store_stage + ew * 2048;
    float* xch_mine = xch + ew * 128;
    const float* xch_peer = xch + (half == 0 ? ew + 4 : ew - 4) * 128;
    const int r = q * 32 + lane;
    const bool row_ok = r < p.Lq;
    int oc = 0;
    for (int i = 0; i < n_mine; ++i) {
      const int u = (int)blockIdx.x + i * (int)gridDim.x;
      const int h = u % p.heads, b = u / p.heads;
      const __nv_bfloat16* prow = p.P + (static_cast<long long>(u) * p.Lq + r) * p.p_ld;
      mbar_wait(s_full, static_cast<uint32_t>(i) & 1u);
      tc_fence_after();
      const uint32_t ts = tmem_base + kColS + (static_cast<uint32_t>(q * 32) << 16);
      // pass 1: this warp's share of sum_c dP P; P is fetched once and kept packed
      float part = 0.f;
      uint4 pk[CPW][4];
#pragma unroll
      for (int cc = 0; cc < CPW; ++cc) {
        const int c = half * CPW + cc;
        uint32_t ra[32];
        tmem_ld32(ts + c * 32, ra);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          // 16-byte loads issued back to back (the row is padded to p_ld, a multiple of 8): the ragged tail of the
          // last group is masked below, columns the forward never wrote are never used
          const int col0 = c * 32 + g * 8;
          pk[cc][g] = (row_ok && col0 < p.Lc) ? __ldg(reinterpret_cast<const uint4*>(prow + col0))
                                              : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int col0 = c * 32 + g * 8;
          if (col0 < p.Lc && col0 + 8 > p.Lc) {
            __nv_bfloat16* h16 = reinterpret_cast<__nv_bfloat16*>(&pk[cc][g]);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (col0 + j >= p.Lc) h16[j] = __float2bfloat16_rn(0.f);
          }
        }
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&pk[cc][g]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = __bfloat1622float2(hh[j]);
            part = fmaf(p.sign * __uint_as_float(ra[g * 8 + 2 * j]), f.x, part);
            part = fmaf(p.sign * __uint_as_float(ra[g * 8 + 2 * j + 1]), f.y, part);
          }
        }
      }
      xch_mine[lane] = part;
      pair_barrier(q);
      const float tot = part + xch_peer[lane];
      // pass 2: dS = alpha P (sign dP - total) -> shared memory, next to the P tile
#pragma unroll
      for (int cc = 0; cc < CPW; ++cc) {
        const int c = half * CPW + cc;
        uint32_t ra[32];
        tmem_ld32(ts + c * 32, ra);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&pk[cc][g]);
          float t8[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = __bfloat1622float2(hh[j]);
            t8[2 * j] = p.alpha * f.x * (p.sign * __uint_as_float(ra[g * 8 + 2 * j]) - tot);
            t8[2 * j + 1] = p.alpha * f.y * (p.sign * __uint_as_float(ra[g * 8 + 2 * j + 1]) - tot);
          }
          const int col0 = c * 32 + g * 8;
          const int off = (col0 >> 6) * kQTile + r * 128 + ((((col0 & 63) >> 3) ^ (r & 7)) << 4);
          *reinterpret_cast<uint4*>(dstile + off) = pack8_bf16(t8);
          *reinterpret_cast<uint4*>(ptile + off) = pk[cc][g];
        }
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(s_free);
        mbar_arrive(ds_ready);
      }
      // ---- output tiles: dV, dQ, dK
      for (int o = 0; o < 3; ++o) {
        const CUtensorMap* tm = o == 0 ? &tmDV : (o == 1 ? &tmDQ : &tmDK);
        const float scale = o == 0 ? p.sign : 1.f;
        const int nco = p.nt >> 5, cpo = nco >> 1;
        for (int t = 0; t < p.n_tiles; ++t) {
          const int ob = oc & 1;
          mbar_wait(&o_full[ob], static_cast<uint32_t>(oc >> 1) & 1u);
          tc_fence_after();
          for (int cc = 0; cc < cpo; ++cc) {
            const int c = half * cpo + cc;
            const int col0 = t * p.nt + c * 32;
            if (col0 >= p.hd) continue;
            uint32_t ra[32];
            float v[32];
            tmem_ld32(tmem_base + kColO + static_cast<uint32_t>(ob * 128) + (static_cast<uint32_t>(q * 32) << 16) + c * 32, ra);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = scale * __uint_as_float(ra[j]);
            tma_store_row32<__nv_bfloat16>(tm, stg, lane, v, col0, q * 32, h, b);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&o_empty[ob]);
          ++oc;
        }
      }
    }
    if (lane == 0) bulk_wait_all();
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int BNS>
int launch_attn_bwd(const CUtensorMap& tmDO, const CUtensorMap& tmV, const CUtensorMap& tmK, const CUtensorMap& tmQ,
                    const CUtensorMap& tmDV, const CUtensorMap& tmDQ, const CUtensorMap& tmDK, const AbParams& p,
                    cudaStream_t stream) {
  using Cfg = AbCfg<BNS>;
  auto kern = attn_bwd_kernel<BNS>;
  static std::atomic<int> attr_set[kMaxDevices];
  const int dev = current_device();
  if (!attr_set[dev].load(std::memory_order_acquire)) {
    D2R_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set[dev].store(1, std::memory_order_release);
  }
  const int grid = p.units < num_sms() ? p.units : num_sms();
  D2R_CUDA_OK(launch_pdl(kern, dim3((unsigned)grid), Cfg::SMEM_BYTES, stream, tmDO, tmV, tmK, tmQ, tmDV, tmDQ, tmDK, p));
  count_launch();
  return check_launch("attn_bwd_kernel");
}

inline bool aligned16(const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; }

}  // namespace

extern "C" {

int d2r_attn_fwd(const d2r_attn_args* a, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  D2R_CHECK_ARG(a != nullptr && a->q && a->k && a->v && a->p && a->out, "attn_fwd: null argument");
  D2R_CHECK_ARG(a->B > 0 && a->heads > 0 && a->Lq > 0 && a->Lc > 0 && a->hd > 0, "attn_fwd: empty problem");
  D2R_CHECK_ARG(a->Lc <= 128, "attn_fwd: at most 128 keys (got %d): use the composed d2r_gemm path", a->Lc);
  D2R_CHECK_ARG(a->hd % 16 == 0 && (a->hd <= 64 || a->hd % 64 == 0), "attn_fwd: head dim %d unsupported", a->hd);
  D2R_CHECK_ARG(a->q_ld % 8 == 0 && a->k_ld % 8 == 0 && a->v_ld % 8 == 0 && a->p_ld % 8 == 0 && a->o_ld % 8 == 0 &&
                    a->hd % 8 == 0 && a->p_ld >= a->Lc,
                "attn_fwd: leading dimensions must be multiples of 8 elements (TMA 16-byte rule)");
  D2R_CHECK_ARG(aligned16(a->q) && aligned16(a->k) && aligned16(a->v) && aligned16(a->p) && aligned16(a->out),
                "attn_fwd: operands must be 16-byte aligned");
  D2R_CHECK_ARG(a->mode == 0 || (a->mode == 1 && a->residual && a->out2 && aligned16(a->out2)),
                "attn_fwd: mode 1 (squared difference) needs residual and out2");
  D2R_CHECK_ARG(!a->residual || (a->r_ld % 8 == 0 && aligned16(a->residual)), "attn_fwd: residual layout");
  const int bns = a->Lc <= 64 ? 64 : 128;
  AfParams p;
  p.Lq = a->Lq; p.Lc = a->Lc; p.hd = a->hd; p.heads = a->heads;
  p.q_tiles = (a->Lq + 127) / 128;
  p.units = a->B * a->heads * p.q_tiles;
  p.nkb = (a->hd + 63) / 64;
  p.ksteps2 = (a->Lc + 15) / 16;
  p.nt = a->hd <= 64 ? 64 : 128;
  p.n_tiles = (a->hd + p.nt - 1) / p.nt;
  p.sc = a->alpha * 1.4426950408889634f;
  p.mode = a->mode;
  p.residual = static_cast<const __nv_bfloat16*>(a->residual);
  p.r_ld = a->r_ld; p.r_sb = (long long)a->Lq * a->r_ld; p.r_sh = a->hd;
  const long long H = a->heads, B = a->B;
  CUtensorMap tmQ, tmK, tmV, tmP, tmO, tmO2;
  memset(&tmO2, 0, sizeof(tmO2));
  int rc;
  // operands: {head dim, rows, heads, samples}; head stride = hd elements, sample stride = rows * ld
  if ((rc = encode_map(&tmQ, a->q, 2, a->hd, a->Lq, H, B, a->q_ld, a->hd, (long long)a->Lq * a->q_ld, 64, 128,
                       CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = encode_map(&tmK, a->k, 2, a->hd, a->Lc, H, B, a->k_ld, a->hd, (long long)a->Lc * a->k_ld, 64, bns,
                       CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = encode_map(&tmV, a->v, 2, a->hd, a->Lc, H, B, a->v_ld, a->hd, (long long)a->Lc * a->v_ld, 64, 64,
                       CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  // P [B, heads, Lq, p_ld]: stored straight from the swizzled shared-memory tile, 32 rows x 64 columns per store
  if ((rc = encode_map(&tmP, a->p, 2, a->Lc, a->Lq, H, B, a->p_ld, (long long)a->Lq * a->p_ld,
                       H * a->Lq * a->p_ld, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  // outputs: 64 columns x 32 rows per store, straight from a 128B-swizzled staging tile
  if ((rc = encode_map(&tmO, a->out, 2, a->hd, a->Lq, H, B, a->o_ld, a->hd, (long long)a->Lq * a->o_ld, 64, 32,
                       CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if (a->mode == 1 &&
      (rc = encode_map(&tmO2, a->out2, 2, a->hd, a->Lq, H, B, a->o_ld, a->hd, (long long)a->Lq * a->o_ld, 64, 32,
                       CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  const bool split = p.nt == 64;          // small heads: one 64-column output tile per unit
  if (bns == 64) return split ? launch_attn_fwd<64, true>(tmQ, tmK, tmV, tmP, tmO, tmO2, p, st)
                              : launch_attn_fwd<64, false>(tmQ, tmK, tmV, tmP, tmO, tmO2, p, st);
  return split ? launch_attn_fwd<128, true>(tmQ, tmK, tmV, tmP, tmO, tmO2, p, st)
               : launch_attn_fwd<128, false>(tmQ, tmK, tmV, tmP, tmO, tmO2, p, st);
}

int d2r_attn_bwd(const d2r_attn_bwd_args* a, void* stream) {
  auto st = static_cast<cudaStream_t>(stream);
  D2R_CHECK_ARG(a != nullptr && a->d_out && a->p && a->q && a->k && a->v && a->dq && a->dk && a->dv,
                "attn_bwd: null argument");
  D2R_CHECK_ARG(a->B > 0 && a->heads > 0 && a->Lq > 0 && a->Lc > 0 && a->hd > 0, "attn_bwd: empty problem");
  D2R_CHECK_ARG(a->Lc <= 128 && a->Lq <= 128,
                "attn_bwd: at most 128 queries and keys per (sample, head) (got %d, %d): use the composed path",
                a->Lq, a->Lc);
  D2R_CHECK_ARG(a->hd % 16 == 0 && (a->hd <= 64 || a->hd % 64 == 0), "attn_bwd: head dim %d unsupported", a->hd);
  D2R_CHECK_ARG(a->do_ld % 8 == 0 && a->q_ld % 8 == 0 && a->k_ld % 8 == 0 && a->v_ld % 8 == 0 && a->p_ld % 8 == 0 &&
                    a->dq_ld % 8 == 0 && a->dk_ld % 8 == 0 && a->dv_ld % 8 == 0 && a->p_ld >= a->Lc,
                "attn_bwd: leading dimensions must be multiples of 8 elements (TMA 16-byte rule)");
  D2R_CHECK_ARG(aligned16(a->d_out) && aligned16(a->p) && aligned16(a->q) && aligned16(a->k) && aligned16(a->v) &&
                    aligned16(a->dq) && aligned16(a->dk) && aligned16(a->dv),
                "attn_bwd: operands must be 16-byte aligned");
  const int bns = a->Lc <= 64 ? 64 : 128;
  AbParams p;
  p.Lq = a->Lq; p.Lc = a->Lc; p.hd = a->hd; p.heads = a->heads;
  p.units = a->B * a->heads;
  p.nkb = (a->hd + 63) / 64;
  p.ksteps_q = (a->Lq + 15) / 16;
  p.ksteps_c = (a->Lc + 15) / 16;
  p.nt = a->hd <= 64 ? 64 : 128;
  p.n_tiles = (a->hd + p.nt - 1) / p.nt;
  p.alpha = a->alpha; p.sign = a->sign;
  p.P = static_cast<const __nv_bfloat16*>(a->p);
  p.p_ld = a->p_ld;
  const long long H = a->heads, B = a->B;
  CUtensorMap tmDO, tmV, tmK, tmQ, tmDV, tmDQ, tmDK;
  int rc;
  auto operand = [&](CUtensorMap* tm, const void* base, long long rows, long long ld, int box_rows) {
    return encode_map(tm, base, 2, a->hd, rows, H, B, ld, a->hd, rows * ld, 64, box_rows, CU_TENSOR_MAP_SWIZZLE_128B);
  };
  auto output = [&](CUtensorMap* tm, void* base, long long rows, long long ld) {
    return encode_map(tm, base, 2, a->hd, rows, H, B, ld, a->hd, rows * ld, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
  };
  if ((rc = operand(&tmDO, a->d_out, a->Lq, a->do_ld, 128))) return rc;
  if ((rc = operand(&tmV, a->v, a->Lc, a->v_ld, bns))) return rc;
  if ((rc = operand(&tmK, a->k, a->Lc, a->k_ld, 128))) return rc;
  if ((rc = operand(&tmQ, a->q, a->Lq, a->q_ld, 128))) return rc;
  if ((rc = output(&tmDV, a->dv, a->Lc, a->dv_ld))) return rc;
  if ((rc = output(&tmDQ, a->dq, a->Lq, a->dq_ld))) return rc;
  if ((rc = output(&tmDK, a->dk, a->Lc, a->dk_ld))) return rc;
  if (bns == 64) return launch_attn_bwd<64>(tmDO, tmV, tmK, tmQ, tmDV, tmDQ, tmDK, p, st);
  return launch_attn_bwd<128>(tmDO, tmV, tmK, tmQ, tmDV, tmDQ, tmDK, p, st);
}

}  // extern "C"
}  // namespace d2r
