// CTA-pair variant of the tcgen05 GEMM (included by gemm_tc.cu inside namespace d2r::<anon>).
//
// Two CTAs of a cluster (one TPC) cooperate on a 256 x 256 output tile with tcgen05.mma.cta_group::2:
// each CTA stages its own 128 rows of A and ITS HALF (128 of 256 N-rows) of B, so the shared-memory fill per
// CTA and k-block drops from 48 KB (128x256 single-CTA tile) to 32 KB for the same MMA work -- the K=768
// projections of this workload are bound by exactly that L2->SM fill rate.  The even CTA (leader) issues the
// MMAs; accumulators live in both CTAs' TMEM (128 lanes each); every CTA runs its own epilogue on its rows.
//
// Synchronisation (all mbarriers at identical smem offsets in both CTAs):
//   full[s]        leader only, count 1 + 64 KB of TMA bytes: both CTAs' loads signal the LEADER's barrier
//   empty[s]       both CTAs, arrived by tcgen05.commit.cta_group::2 ... multicast (mask 0b11)
//   tmem_full[a]   both CTAs, multicast commit after the last k-block of a tile
//   tmem_empty[a]  leader only, count 16: 8 local epilogue warps + 8 remote arrives (mapa) from the peer
//
// PAIRS == 2 ("quad"): a cluster of two pairs stacked along M works on a 512 x 256 tile.  Both pairs need the same
// B tile, so each CTA fetches only a 64-row quarter of its B half and TMA-multicasts it to its counterpart in the
// other pair: the L2->SM fill per CTA and k-block drops again, 32 KB -> 24 KB.  A stage may then only be refilled
// when BOTH pairs have consumed it: empty[s] counts one multicast commit per pair, sent to all four CTAs.
#pragma once

struct Tc2Cfg {
  static constexpr int BN = 256;
  static constexpr int HALF_N = 128;
  static constexpr int B_STAGE_BYTES = HALF_N * BK * 2;               // 16 KB
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;   // 32 KB per CTA
  static constexpr int STAGES = 6;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int BAR_BYTES = 256;
  static constexpr int BIAS_BYTES = 8 * BN * 4;
  static constexpr int STORE_BYTES = 8 * 2048;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STORE_BYTES + BAR_BYTES + BIAS_BYTES + 1024;
};

template <int PAIRS, bool A_MN, bool B_MN>
__global__ void __cluster_dims__(2 * PAIRS, 1, 1) __launch_bounds__(kTcThreads, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2, const TcParams p) {
  using Cfg = Tc2Cfg;
  constexpr int BN = Cfg::BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* store_stage = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(store_stage + Cfg::STORE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full = empty_bar + Cfg::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* sbias = reinterpret_cast<float*>(store_stage + Cfg::STORE_BYTES + Cfg::BAR_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const uint32_t rank = crank & 1;                  // inside the pair: 0 = leader (issues the MMAs), 1 = peer
  const uint32_t pr = crank >> 1;                   // pair index inside the cluster (PAIRS == 2: stacked along M)
  const long long pair = blockIdx.x / (2 * PAIRS);  // cluster index: one BM*2*PAIRS x 256 tile at a time
  const long long npairs = gridDim.x / (2 * PAIRS);
  constexpr uint16_t kAllCtas = PAIRS == 2 ? 0xF : 0x3;
  const bool prof = p.prof != nullptr;
  long long* rec = prof ? p.prof + static_cast<long long>(blockIdx.x) * kProfSlots : nullptr;
  const long long t_entry = prof ? clock64() : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tma_store) {
      tma_prefetch_desc(&tmC);
      if (p.epilogue == D2R_EPI_SQDIFF) tma_prefetch_desc(&tmC2);
    }
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], PAIRS);              // one multicast commit per pair of the cluster
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 16);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                               // both CTAs' barriers are initialised before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();                                    // (see gemm_tc_kernel: the prologue touched no global data)
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer (one per CTA)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      long long w_empty = 0, nkb = 0;
      const long long t_loop = prof ? clock64() : 0;
      for (long long t = pair; t < p.num_tiles; t += npairs) {
        const TileCoord tc = decode_tile(p, t, BN);
        const int m0 = tc.m0 + static_cast<int>(pr) * 2 * BM + static_cast<int>(rank) * BM;
        const int nh = tc.n0 + static_cast<int>(rank) * Cfg::HALF_N;
        const int azi = p.a_bcast_i ? 0 : tc.zi, azo = p.a_bcast_o ? 0 : tc.zo;
        const int bzi = p.b_bcast_i ? 0 : tc.zi, bzo = p.b_bcast_o ? 0 : tc.zo;
        // (an explicit TMA L2 prefetch of the next tile's A rows was measured and made every shape slower:
        //  it competes with the demand loads for the same L2 request bandwidth)
        for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
          mbar_wait_timed(&empty_bar[stage], phase ^ 1, prof, w_empty);
          ++nkb;
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          if constexpr (!A_MN) {
            tma_load_4d_2sm(sa, &tmA, &full_bar[stage], kb * BK, m0, azi, azo);
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i)
              tma_load_4d_2sm(sa + i * ATOM_BYTES, &tmA, &full_bar[stage], m0 + 64 * i, kb * BK, azi, azo);
          }
          if constexpr (PAIRS == 2) {
            // both pairs of the cluster multiply by the same B tile: this CTA fetches one 64-row quarter of its
            // half and multicasts it to the CTA of the same parity in the other pair (and to itself)
            const uint16_t mc = static_cast<uint16_t>(0x5u << rank);
            const int nq = nh + 64 * static_cast<int>(pr);
            uint8_t* sq = sb + pr * ATOM_BYTES;         // 64 rows x 128 B (K-major) or one 64x64 atom (MN-major)
            if constexpr (!B_MN) tma_load_4d_2sm_mc(sq, &tmB, &full_bar[stage], mc, kb * BK, nq, bzi, bzo);
            else                 tma_load_4d_2sm_mc(sq, &tmB, &full_bar[stage], mc, nq, kb * BK, bzi, bzo);
          } else if constexpr (!B_MN) {
            tma_load_4d_2sm(sb, &tmB, &full_bar[stage], kb * BK, nh, bzi, bzo);
          } else {
#pragma unroll
            for (int i = 0; i < Cfg::HALF_N / 64; ++i)
              tma_load_4d_2sm(sb + i * ATOM_BYTES, &tmB, &full_bar[stage], nh + 64 * i, kb * BK, bzi, bzo);
          }
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      if (prof) {
        rec[PS_PROD_WAIT_EMPTY] = w_empty;
        rec[PS_PROD_KBLOCKS] = nkb;
        rec[PS_PROD_LOOP] = clock64() - t_loop;
        if (rank != 0) {
          rec[PS_START] = t_entry;
          rec[PS_SMID] = smid();
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer (leader CTA only)
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      const uint16_t my_pair = static_cast<uint16_t>(0x3u << (2 * pr));
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      long long w_full = 0, w_tmem = 0, ntiles = 0, first_full = -1;
      const long long t_loop = prof ? clock64() : 0;
      for (long long t = pair; t < p.num_tiles; t += npairs) {
        const TileCoord tc = decode_tile(p, t, BN);
        mbar_wait_timed(&tmem_empty[acc], acc_phase ^ 1, prof, w_tmem);
        ++ntiles;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
          mbar_wait_timed(&full_bar[stage], phase, prof, w_full);
          if (prof && first_full < 0) first_full = clock64() - t_loop;
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da = A_MN ? make_smem_desc_sw128(sa + k * 2048, ATOM_BYTES, 1024)
                                     : make_smem_desc_sw128(sa + k * 32, 16, 1024);
            const uint64_t db = B_MN ? make_smem_desc_sw128(sb + k * 2048, ATOM_BYTES, 1024)
                                     : make_smem_desc_sw128(sb + k * 32, 16, 1024);
            umma_bf16_2sm(d_tmem, da, db, idesc, (kb > tc.kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit_2sm(&empty_bar[stage], kAllCtas); // frees the slot in every CTA that (multi)casts into it
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2sm(&tmem_full[acc], my_pair);      // accumulator complete -> both epilogues of this pair
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      if (prof) {
        const long long now = clock64();
        rec[PS_START] = t_entry;
        rec[PS_END] = now;
        rec[PS_MMA_WAIT_FULL] = w_full;
        rec[PS_MMA_WAIT_TMEM] = w_tmem;
        rec[PS_MMA_LOOP] = now - t_loop;
        rec[PS_TILES] = ntiles;
        rec[PS_PROLOGUE] = t_loop - t_entry;
        rec[PS_FIRST_FULL] = first_full;
        rec[PS_SMID] = smid();
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------- epilogue (8 warps per CTA, own 128 rows)
    const int q = warp & 3;
    float* sb = sbias + (warp - 2) * (BN < 128 ? 128 : BN);
    const int half = (warp - 2) >> 2;
    uint8_t* stg = store_stage + (warp - 2) * 2048;
    const TileWalk walk{pair, npairs, static_cast<int>(pr) * 2 * BM + static_cast<int>(rank) * BM,
                        rank == 0 ? -1 : static_cast<int>(2 * pr)};
    using bf16 = __nv_bfloat16;
    switch (p.variant) {
      case EV_F32:        epilogue_loop<BN, float, NoRes, 0>(p, &tmC, &tmC2, tmem_base, tmem_full, tmem_empty, sb, stg, q, half, lane, walk); break;
      case EV_F32_RF32:   epilogue_loop<BN, float, float, 0>(p, &tmC, &tmC2, tmem_base, tmem_full, tmem_empty, sb, stg, q, half, lane, walk); break;
      case EV_F32_RBF16:  epilogue_loop<BN, float, bf16, 0>(p, &tmC, &tmC2, tmem_base, tmem_full, tmem_empty, sb, stg, q, half, lane, walk); break;
      case EV_BF16:       epilogue_loop<BN, bf16, NoRes, 0>(p, &tmC, &tmC2, tmem_base, tmem_full, tmem_empty, sb, stg, q, half, lane, walk); break;
      case EV_BF16_RF32:  epilogue_loop<BN, bf16, float, 0>(p, &tmC, &tmC2, tmem_base, tmem_full, tmem_empty, sb, stg, q, half, lane, walk); break;
      case EV_BF16_RBF16: epilogue_loop<BN, bf16, bf16, 0>(p, &tmC, &tmC2, tmem_base, tmem_full, tmem_empty, sb, stg, q, half, lane, walk); break;
      case EV_SQ_F32:     epilogue_loop<BN, float, float, 1>(p, &tmC, &tmC2, tmem_base, tmem_full, tmem_empty, sb, stg, q, half, lane, walk); break;
      case EV_SQ_BF16:    epilogue_loop<BN, bf16, bf16, 1>(p, &tmC, &tmC2, tmem_base, tmem_full, tmem_empty, sb, stg, q, half, lane, walk); break;
      case EV_SOFTMAX_BF16:     epilogue_loop<BN, bf16, NoRes, 3>(p, &tmC, &tmC2, tmem_base, tmem_full, tmem_empty, sb, stg, q, half, lane, walk); break;
      case EV_SOFTMAX_BWD_BF16: epilogue_loop<BN, bf16, bf16, 4>(p, &tmC, &tmC2, tmem_base, tmem_full, tmem_empty, sb, stg, q, half, lane, walk); break;
      default:            epilogue_loop<BN, float, NoRes, 2>(p, &tmC, &tmC2, tmem_base, tmem_full, tmem_empty, sb, stg, q, half, lane, walk); break;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();           // neither CTA leaves (or frees TMEM) while its partner may still signal it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
  }
}
