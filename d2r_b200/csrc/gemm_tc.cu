// tcgen05 / TMEM / TMA batched GEMM with fused epilogue (bf16 operands, fp32 accumulation).
//
//   C[z] = epilogue(alpha * A[z] (m x k) * B[z]^T (n x k))
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0      TMA producer   (cp.async.bulk.tensor.4d -> 128B-swizzled smem ring)
//   warp 1      MMA issuer     (one thread: tcgen05.mma cta_group::1 kind::f16, M=128, N=BN, K=16)
//   warps 2..9  epilogue       (tcgen05.ld 32x32b -> registers -> bias/act/residual -> global)
// Accumulators are double-buffered in TMEM (2 x BN fp32 columns) so the epilogue of tile i
// overlaps the main loop of tile i+1.  Operands may be K-major or MN-major (transposed) so the
// same kernel serves forward (x W^T), data-gradient (dy W) and weight-gradient (dy^T x).
#include <cuda.h>

#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"

namespace d2r {

// debug probe (d2r_gemm_set_profile): device buffer of per-CTA stall records, nullptr = profiling off
std::atomic<long long*> g_prof_buf{nullptr};

namespace {

#include "tc_host.cuh"

constexpr int BM = 128;
constexpr int BK = 64;            // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int ATOM_BYTES = 64 * BK * 2;   // one [64 x 64] bf16 swizzle tile = 8 KB

// Accumulator buffers in TMEM (512 fp32 columns per SM).  Wide tiles are double-buffered; the narrow tiles serve the
// small-K batched products of the attention blocks, where a tile's main loop is one or two k-blocks and the tile
// rate is set by the MMA -> epilogue -> MMA barrier round trip: more buffers keep more tiles in flight.
constexpr int num_acc(int bn) { return bn >= 192 ? 2 : (bn == 128 ? 4 : 8); }

template <int BN>
struct TcCfg {
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = BN == 256 ? 4 : (BN == 192 ? 5 : (BN == 128 ? 6 : 8));
  static constexpr int NACC = num_acc(BN);
  static constexpr int TMEM_COLS = 512;            // accumulator a at column a * BN
  static constexpr int BAR_BYTES = 512;
  static constexpr int BIAS_BYTES = 8 * (BN < 128 ? 128 : BN) * 4;   // one private bias slice per epilogue warp
                                                                      // (>= 128 floats: softmax exchange slots)
  static constexpr int STORE_BYTES = 8 * 2048;    // one 32-row x 64-byte staging tile per epilogue warp
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STORE_BYTES + BAR_BYTES + BIAS_BYTES + 1024;
};

struct TcParams {
  int m, n, k;
  int bm;            // rows per tile: 128 (one CTA) or 256 (CTA pair, cta_group::2)
  int batch_inner;
  int m_tiles, n_tiles, k_blocks, split_k, kb_per_split;
  long long num_tiles;
  int a_bcast_i, a_bcast_o, b_bcast_i, b_bcast_o;   // 1: batch stride 0 -> coordinate pinned to 0
  void* c;
  void* c2;
  const float* bias;
  const void* residual;
  long long ldc, ldr, c_so, c_si, r_so, r_si, bias_sz;
  float alpha;
  int act, epilogue, c_dtype, r_dtype, atomic, act_cols, variant;
  int tma_store;   // 1: C (and c2) are written with TMA bulk tensor stores through a smem staging tile
  int debug;   // D2R_TC_DEBUG env (bring-up only): 1 = skip epilogue stores, 2 = skip the epilogue body, 4 = no TMA stores
  long long* prof;   // d2r_gemm_set_profile: per-CTA stall record (kProfSlots int64 each), nullptr = off
};

// Per-CTA stall record written when p.prof is set (tools/gemm_stall.py reads it).  All values are SM clock cycles
// of ONE elected thread per role: where the TMA producer, the MMA issuer and one epilogue warp spend their time.
constexpr int kProfSlots = 16;
enum ProfSlot {
  PS_START = 0, PS_END, PS_PROD_WAIT_EMPTY, PS_PROD_KBLOCKS, PS_PROD_LOOP, PS_MMA_WAIT_FULL, PS_MMA_WAIT_TMEM,
  PS_MMA_LOOP, PS_TILES, PS_EPI_WAIT_FULL, PS_EPI_LOOP, PS_PROLOGUE, PS_FIRST_FULL, PS_SMID, PS_EPI_STORE_WAIT
};
__device__ __forceinline__ void mbar_wait_timed(uint64_t* bar, uint32_t parity, bool on, long long& acc) {
  if (on) {
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    acc += clock64() - t0;
  } else {
    mbar_wait(bar, parity);
  }
}
__device__ __forceinline__ uint32_t smid() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
  return r;
}

struct TileCoord {
  int m0, n0, zi, zo, z, kb0, kb1;
};

__device__ __forceinline__ TileCoord decode_tile(const TcParams& p, long long t, int bn) {
  TileCoord tc;
  // 32-bit arithmetic (the host rejects launches with >= 2^31 tiles): a 64-bit divide is ~100 instructions, and
  // the three role warps each decode every tile
  unsigned u = static_cast<unsigned>(t);
  const int nb = static_cast<int>(u % static_cast<unsigned>(p.n_tiles));
  u /= static_cast<unsigned>(p.n_tiles);
  const int ks = static_cast<int>(u % static_cast<unsigned>(p.split_k));
  u /= static_cast<unsigned>(p.split_k);
  const int mb = static_cast<int>(u % static_cast<unsigned>(p.m_tiles));
  const int z = static_cast<int>(u / static_cast<unsigned>(p.m_tiles));
  tc.m0 = mb * p.bm;
  tc.n0 = nb * bn;
  tc.z = z;
  tc.zi = z % p.batch_inner;
  tc.zo = z / p.batch_inner;
  tc.kb0 = ks * p.kb_per_split;
  tc.kb1 = min(p.k_blocks, tc.kb0 + p.kb_per_split);
  return tc;
}

#include "gemm_tc_epilogue.cuh"

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2, const TcParams p) {
  using Cfg = TcCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled tiles must start on a 1024-byte boundary
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* store_stage = smem + Cfg::STAGES * Cfg::STAGE_BYTES;   // [8][2048]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(store_stage + Cfg::STORE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full = empty_bar + Cfg::STAGES;
  uint64_t* tmem_empty = tmem_full + Cfg::NACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + Cfg::NACC);
  float* sbias = reinterpret_cast<float*>(store_stage + Cfg::STORE_BYTES + Cfg::BAR_BYTES);   // [8][BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
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
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < Cfg::NACC; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // The prologue above touches no global data: under programmatic dependent launch it overlaps the tail of the
  // previous kernel on the stream.  From here on every role reads or overwrites memory that kernel may still use.
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      long long w_empty = 0, nkb = 0;
      const long long t_loop = prof ? clock64() : 0;
      for (long long t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const TileCoord tc = decode_tile(p, t, BN);
        const int azi = p.a_bcast_i ? 0 : tc.zi, azo = p.a_bcast_o ? 0 : tc.zo;
        const int bzi = p.b_bcast_i ? 0 : tc.zi, bzo = p.b_bcast_o ? 0 : tc.zo;
        for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
          mbar_wait_timed(&empty_bar[stage], phase ^ 1, prof, w_empty);
          ++nkb;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          if constexpr (!A_MN) {
            tma_load_4d(sa, &tmA, &full_bar[stage], kb * BK, tc.m0, azi, azo);
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i)
              tma_load_4d(sa + i * ATOM_BYTES, &tmA, &full_bar[stage], tc.m0 + 64 * i, kb * BK, azi, azo);
          }
          if constexpr (!B_MN) {
            tma_load_4d(sb, &tmB, &full_bar[stage], kb * BK, tc.n0, bzi, bzo);
          } else {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i)
              tma_load_4d(sb + i * ATOM_BYTES, &tmB, &full_bar[stage], tc.n0 + 64 * i, kb * BK, bzi, bzo);
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
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      long long w_full = 0, w_tmem = 0, ntiles = 0, first_full = -1;
      const long long t_loop = prof ? clock64() : 0;
      for (long long t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
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
            // K-major: 8-row groups are 1024 B apart (SBO), a K step of 16 elements is 32 B.
            // MN-major: 64-wide MN tiles are 8 KB apart (LBO), 8-deep K groups 1024 B apart (SBO),
            //           a K step of 16 is two K groups = 2048 B.
            const uint64_t da = A_MN ? make_smem_desc_sw128(sa + k * 2048, ATOM_BYTES, 1024)
                                     : make_smem_desc_sw128(sa + k * 32, 16, 1024);
            const uint64_t db = B_MN ? make_smem_desc_sw128(sb + k * 2048, ATOM_BYTES, 1024)
                                     : make_smem_desc_sw128(sb + k * 32, 16, 1024);
            umma_bf16(d_tmem, da, db, idesc, (kb > tc.kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);   // frees the smem slot once these MMAs retire
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);        // accumulator complete -> epilogue
        if (++acc == Cfg::NACC) {
          acc = 0;
          acc_phase ^= 1;
        }
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
    // ------------------------------------------------------------- epilogue (4 warps)
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    float* sb = sbias + (warp - 2) * (BN < 128 ? 128 : BN);
    const int half = (warp - 2) >> 2;
    uint8_t* stg = store_stage + (warp - 2) * 2048;
    const TileWalk walk{static_cast<long long>(blockIdx.x), static_cast<long long>(gridDim.x), 0, -1};
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
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

#include "gemm_tc2.cuh"

template <int BN, bool A_MN, bool B_MN>
int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmC2,
              const TcParams& p, cudaStream_t stream) {
  using Cfg = TcCfg<BN>;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN>;
  // per device, and safe against the forward thread and autograd's backward thread racing on the first launch
  // (setting the attribute twice is harmless; the flag is only ever raised after a successful call)
  static std::atomic<int> attr_set[kMaxDevices];
  const int dev = current_device();
  if (!attr_set[dev].load(std::memory_order_acquire)) {
    D2R_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set[dev].store(1, std::memory_order_release);
  }
  long long grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  D2R_CUDA_OK(launch_pdl(kern, dim3((unsigned)grid), Cfg::SMEM_BYTES, stream, tmA, tmB, tmC, tmC2, p));
  count_launch();
  return check_launch("gemm_tc_kernel");
}

// How many clusters of `ctas` CTAs (1 CTA per SM) the device can hold at once: GPC boundaries make this less than
// num_sms / ctas for the 4-CTA cluster, and a persistent grid must not be larger than what is co-resident.
template <typename Kern>
int max_clusters(Kern kern, int ctas, int smem_bytes) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(num_sms() / ctas * ctas), 1, 1);
  cfg.blockDim = dim3(kTcThreads, 1, 1);
  cfg.dynamicSmemBytes = smem_bytes;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    n = num_sms() / ctas;
  }
  return n;
}

template <int PAIRS, bool A_MN, bool B_MN>
int launch_tc2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmC2,
               const TcParams& p, cudaStream_t stream) {
  auto kern = gemm_tc2_kernel<PAIRS, A_MN, B_MN>;
  static std::atomic<long long> max_units_dev[kMaxDevices];   // per device; 0 = not initialised yet (see launch_tc)
  const int dev = current_device();
  long long max_units = max_units_dev[dev].load(std::memory_order_acquire);
  if (!max_units) {
    D2R_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Tc2Cfg::SMEM_BYTES));
    max_units = PAIRS == 1 ? num_sms() / 2 : max_clusters(kern, 2 * PAIRS, Tc2Cfg::SMEM_BYTES);
    max_units_dev[dev].store(max_units, std::memory_order_release);
  }
  const long long units = p.num_tiles < max_units ? p.num_tiles : max_units;
  D2R_CUDA_OK(launch_pdl(kern, dim3((unsigned)(2 * PAIRS * units)), Tc2Cfg::SMEM_BYTES, stream, tmA, tmB, tmC, tmC2, p));
  count_launch();
  return check_launch("gemm_tc2_kernel");
}

template <int PAIRS>
int launch_tc2_major(bool a_mn, bool b_mn, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                     const CUtensorMap& tmC2, const TcParams& p, cudaStream_t stream) {
  if (!a_mn && !b_mn) return launch_tc2<PAIRS, false, false>(tmA, tmB, tmC, tmC2, p, stream);
  if (!a_mn && b_mn) return launch_tc2<PAIRS, false, true>(tmA, tmB, tmC, tmC2, p, stream);
  if (a_mn && !b_mn) return launch_tc2<PAIRS, true, false>(tmA, tmB, tmC, tmC2, p, stream);
  return launch_tc2<PAIRS, true, true>(tmA, tmB, tmC, tmC2, p, stream);
}

template <int BN>
int launch_tc_major(bool a_mn, bool b_mn, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                    const CUtensorMap& tmC2, const TcParams& p, cudaStream_t stream) {
  if (!a_mn && !b_mn) return launch_tc<BN, false, false>(tmA, tmB, tmC, tmC2, p, stream);
  if (!a_mn && b_mn) return launch_tc<BN, false, true>(tmA, tmB, tmC, tmC2, p, stream);
  if (a_mn && !b_mn) return launch_tc<BN, true, false>(tmA, tmB, tmC, tmC2, p, stream);
  return launch_tc<BN, true, true>(tmA, tmB, tmC, tmC2, p, stream);
}

}  // namespace

int gemm_tc_prof_slots() { return kProfSlots; }

int gemm_tc(const d2r_gemm_args& a, cudaStream_t stream) {
  D2R_CHECK_ARG(a.m > 0 && a.n > 0 && a.k > 0 && a.batch > 0 && a.batch_inner > 0, "gemm: empty problem");
  D2R_CHECK_ARG(a.batch % a.batch_inner == 0, "gemm: batch %d not a multiple of batch_inner %d", a.batch,
                a.batch_inner);
  D2R_CHECK_ARG((reinterpret_cast<uintptr_t>(a.a) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.b) & 15) == 0,
                "gemm(bf16): A and B must be 16-byte aligned");
  D2R_CHECK_ARG(a.lda % 8 == 0 && a.ldb % 8 == 0 && a.a_so % 8 == 0 && a.a_si % 8 == 0 && a.b_so % 8 == 0 &&
                    a.b_si % 8 == 0,
                "gemm(bf16): lda/ldb/batch strides must be multiples of 8 elements (TMA 16-byte rule)");
  D2R_CHECK_ARG(a.epilogue != D2R_EPI_SQDIFF || (a.residual && a.c2), "gemm: SQDIFF needs residual and c2");
  const bool softmax_epi = a.epilogue == D2R_EPI_SOFTMAX || a.epilogue == D2R_EPI_SOFTMAX_BWD;
  if (softmax_epi) {
    D2R_CHECK_ARG(a.n <= 256 && a.split_k <= 1 && !a.accumulate && a.act == D2R_ACT_NONE && !a.bias &&
                      a.c_dtype == D2R_BF16,
                  "gemm: the fused softmax epilogues need n <= 256 (one N tile), bf16 output, no bias/act/split-K");
    D2R_CHECK_ARG(a.epilogue != D2R_EPI_SOFTMAX_BWD || (a.residual && a.r_dtype == D2R_BF16),
                  "gemm: SOFTMAX_BWD needs the bf16 probabilities as residual");
  }
  const int bo = a.batch / a.batch_inner, bi = a.batch_inner;
  int split_k = a.split_k > 1 ? a.split_k : 1;

  int bn = a.tile_n;
  bool pair = false;                       // CTA-pair kernel: 256 x 256 tiles, tcgen05.mma.cta_group::2
  bool quad = false;                       // two pairs per cluster, 512 x 256 tiles, B multicast across the pairs
  if (bn == 512 || bn == 1024) {
    pair = true;
    quad = bn == 1024;
    bn = 256;
  } else if (bn == 0) {
    bn = a.n <= 64 ? 64 : (a.n <= 128 ? 128 : 256);
    if (!softmax_epi) {
      // latency-bound launches (the [B,768] global branches: 6 tiles of 128x256): narrower tiles put more SMs on
      // the same work -- measured 16.4 -> 12.3 us for m=256, n=768, k=768
      const long long mt = (a.m + BM - 1) / BM;
      while (bn > 64 && mt * ((a.n + bn - 1) / bn) * split_k * a.batch < 32) bn >>= 1;
    }
    if (bn == 256) {
      const long long work = 1LL * ((a.m + 255) / 256) * ((a.n + 255) / 256) * split_k * a.batch;
      // enough 256x256 tiles to occupy the 74 CTA pairs, and a shape where halving the B fill pays (measured:
      // long k, wide n or tall m -- 12800-row K=768 projection: 25.6 us on pairs, 27.5 us on single-CTA tiles)
      pair = !softmax_epi && work >= 32 && (a.k >= 1536 || a.n >= 1536 || a.m >= 8192);
    }
  }
  D2R_CHECK_ARG(bn == 64 || bn == 128 || bn == 192 || bn == 256, "gemm: tile_n %d unsupported", bn);
  D2R_CHECK_ARG(!softmax_epi || (!pair && bn >= a.n), "gemm: the softmax epilogues need the whole row in one tile");
  const int bm = quad ? 4 * BM : (pair ? 2 * BM : BM);

  TcParams p;
  p.m = a.m; p.n = a.n; p.k = a.k;
  p.bm = bm;
  p.batch_inner = bi;
  p.m_tiles = (a.m + bm - 1) / bm;
  p.n_tiles = (a.n + bn - 1) / bn;
  p.k_blocks = (a.k + BK - 1) / BK;
  if (split_k > p.k_blocks) split_k = p.k_blocks;
  p.kb_per_split = (p.k_blocks + split_k - 1) / split_k;
  p.split_k = (p.k_blocks + p.kb_per_split - 1) / p.kb_per_split;
  p.num_tiles = 1LL * p.m_tiles * p.n_tiles * p.split_k * a.batch;
  D2R_CHECK_ARG(p.num_tiles < (1LL << 31), "gemm: too many tiles");
  const bool atomic = a.accumulate || p.split_k > 1;
  D2R_CHECK_ARG(!atomic || a.c_dtype == D2R_F32, "gemm: accumulate/split_k need an fp32 C");
  D2R_CHECK_ARG(!atomic || (a.epilogue == D2R_EPI_STD && a.act == D2R_ACT_NONE && !a.residual),
                "gemm: accumulate/split_k support only the plain epilogue");
  p.a_bcast_i = (bi > 1 && a.a_si == 0); p.a_bcast_o = (bo > 1 && a.a_so == 0);
  p.b_bcast_i = (bi > 1 && a.b_si == 0); p.b_bcast_o = (bo > 1 && a.b_so == 0);
  p.c = a.c; p.c2 = a.c2; p.bias = a.bias; p.residual = a.residual;
  p.ldc = a.ldc; p.ldr = a.ldr; p.c_so = a.c_so; p.c_si = a.c_si; p.r_so = a.r_so; p.r_si = a.r_si;
  p.bias_sz = a.bias_sz;
  p.alpha = a.alpha; p.act = a.act; p.epilogue = a.epilogue; p.c_dtype = a.c_dtype; p.r_dtype = a.r_dtype;
  p.atomic = atomic ? 1 : 0;
  p.act_cols = a.act_cols;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("D2R_TC_DEBUG"); dbg = e ? atoi(e) : 0; }
    p.debug = dbg;
  }
  p.prof = g_prof_buf.load(std::memory_order_relaxed);
  D2R_CHECK_ARG(a.act_cols % 8 == 0, "gemm: act_cols must be a multiple of 8");
  if (atomic) {
    p.variant = EV_ATOMIC;
  } else if (a.epilogue == D2R_EPI_SOFTMAX) {
    p.variant = EV_SOFTMAX_BF16;
  } else if (a.epilogue == D2R_EPI_SOFTMAX_BWD) {
    p.variant = EV_SOFTMAX_BWD_BF16;
  } else if (a.epilogue == D2R_EPI_SQDIFF) {
    D2R_CHECK_ARG(a.r_dtype == a.c_dtype, "gemm: SQDIFF needs residual and outputs of the same dtype");
    p.variant = a.c_dtype == D2R_BF16 ? EV_SQ_BF16 : EV_SQ_F32;
  } else {
    const int base = a.c_dtype == D2R_BF16 ? EV_BF16 : EV_F32;
    p.variant = base + (a.residual ? (a.r_dtype == D2R_BF16 ? 2 : 1) : 0);
  }

  if (p.split_k > 1 && !a.accumulate) {
    // split-K partial sums are combined with atomics: C must start at zero (dense C only)
    D2R_CHECK_ARG(a.batch == 1 && a.ldc == a.n, "gemm: split_k without accumulate needs a dense, unbatched C");
    D2R_CUDA_OK(cudaMemsetAsync(a.c, 0, sizeof(float) * (size_t)a.m * a.n, stream));
  }

  CUtensorMap tmA, tmB;
  int rc;
  // K-major operand: inner = k, rows = m|n, box rows = tile rows.  MN-major: inner = m|n, rows = k, box rows = BK.
  const long long a_bi = p.a_bcast_i ? 1 : bi, a_bo = p.a_bcast_o ? 1 : bo;
  const long long b_bi = p.b_bcast_i ? 1 : bi, b_bo = p.b_bcast_o ? 1 : bo;
  if (!a.a_mn_major) rc = encode_operand(&tmA, a.a, a.k, a.m, a_bi, a_bo, a.lda, a.a_si, a.a_so, BM);
  else               rc = encode_operand(&tmA, a.a, a.m, a.k, a_bi, a_bo, a.lda, a.a_si, a.a_so, BK);
  if (rc) return rc;
  if (!a.b_mn_major) rc = encode_operand(&tmB, a.b, a.k, a.n, b_bi, b_bo, a.ldb, a.b_si, a.b_so, quad ? 64 : (pair ? 128 : bn));
  else               rc = encode_operand(&tmB, a.b, a.n, a.k, b_bi, b_bo, a.ldb, a.b_si, a.b_so, BK);
  if (rc) return rc;

  // C / c2 through TMA bulk stores when their layout obeys the 16-byte rules (always true for the stack's tensors)
  CUtensorMap tmC, tmC2;
  memset(&tmC, 0, sizeof(tmC));
  memset(&tmC2, 0, sizeof(tmC2));
  {
    const int es = a.c_dtype == D2R_BF16 ? 2 : 4;
    const long long q = 16 / es;
    const bool ok = (reinterpret_cast<uintptr_t>(a.c) & 15) == 0 && a.ldc % q == 0 &&
                    (bi == 1 || (a.c_si % q == 0 && a.c_si != 0)) && (bo == 1 || (a.c_so % q == 0 && a.c_so != 0)) &&
                    (a.epilogue != D2R_EPI_SQDIFF || (reinterpret_cast<uintptr_t>(a.c2) & 15) == 0);
    p.tma_store = (ok && !(p.debug & 4)) ? 1 : 0;   // debug bit 4: force per-thread global stores
    if (ok) {
      rc = encode_map(&tmC, a.c, es, a.n, a.m, bi, bo, a.ldc, a.c_si, a.c_so, 64 / es, 32, CU_TENSOR_MAP_SWIZZLE_64B);
      if (rc) return rc;
      if (a.epilogue == D2R_EPI_SQDIFF) {
        rc = encode_map(&tmC2, a.c2, es, a.n, a.m, bi, bo, a.ldc, a.c_si, a.c_so, 64 / es, 32,
                        CU_TENSOR_MAP_SWIZZLE_64B);
        if (rc) return rc;
      }
    }
  }
  const bool amn = a.a_mn_major != 0, bmn = a.b_mn_major != 0;
  if (quad) return launch_tc2_major<2>(amn, bmn, tmA, tmB, tmC, tmC2, p, stream);
  if (pair) return launch_tc2_major<1>(amn, bmn, tmA, tmB, tmC, tmC2, p, stream);
  if (bn == 64) return launch_tc_major<64>(amn, bmn, tmA, tmB, tmC, tmC2, p, stream);
  if (bn == 128) return launch_tc_major<128>(amn, bmn, tmA, tmB, tmC, tmC2, p, stream);
  if (bn == 192) return launch_tc_major<192>(amn, bmn, tmA, tmB, tmC, tmC2, p, stream);
  return launch_tc_major<256>(amn, bmn, tmA, tmB, tmC, tmC2, p, stream);
}

}  // namespace d2r
