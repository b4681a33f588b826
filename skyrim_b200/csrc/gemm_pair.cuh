// A-stationary GEMM on CTA pairs (tcgen05 cta_group::2) for wide outputs (QKV projection):
//
//     D[M, N] = A[M, K] W[N, K]^T  -> fused epilogue,   K = C (192 | 384),  N = n_tiles * 192
//
// k_gemm2 streams, per 128 x 192 output tile, the A tile AND the weight tile from L2; for the QKV
// projection at C = 384 that is 1.5 GB of L2 -> SM traffic per launch at 11.2 TB/s, the measured
// cap of the L2 fabric (profiles/r1_qkv_pair.md) — the kernel was L2-bandwidth bound at 2.5x its
// DRAM time.  Here a CTA pair owns 256 token rows: each CTA keeps ITS 128-row A tile resident in
// shared memory for all n-tiles and streams only HALF of every weight item (the M = 256 MMA reads
// the two halves from the two CTAs), which cuts the traffic to 0.55 GB per launch.
//
// Per CTA, 320 threads: warps 0..7 epilogue, 8 loader, 9 MMA issuer (even CTA) / completion relay
// (odd CTA; 1-D bulk copies cannot signal the peer's mbarrier).  Accumulators: 2 x 192 TMEM columns,
// so the epilogue of n-tile i overlaps the MMAs of n-tile i+1.  The A tile is released k-block by
// k-block during the last n-tile, and the loader interleaves the next tile's A k-blocks with its
// first weight items in consumption order, so there is no refill bubble between tiles.
#pragma once
#include <cstdlib>

#include "gemm2.cuh"

namespace sky {

template <int C, int NT_ = 192>
struct GPairCfg {
  static constexpr int NKB = C / 64;
  static constexpr int NT = NT_;                 // n-tile = one MMA N (192: Pangu QKV; 256: GraphCast hidden layers, K = 512)
  static constexpr int W_FULL = NT * 128;        // one (n-tile, k-block) item of the weight image
  static constexpr int W_HALF = W_FULL / 2;      // this CTA's NT / 2 rows of it
  static constexpr int S = C == 192 ? 10 : C == 384 ? 7 : 4;
  static_assert(2 * NT <= 512 && NT % 16 == 0 && NT <= 256, "two accumulators in TMEM, one UMMA N");
  static constexpr int A_BYTES = NKB * G2_A_BYTES;
  static constexpr int OFF_W = A_BYTES;
  static constexpr int OFF_PATCH = OFF_W + S * W_HALF;
  static constexpr int OFF_BAR = OFF_PATCH + 8 * G2_PATCH_FLOATS * 4;
  static constexpr int NBARS = 3 * NKB + 3 * S + 4;
  static constexpr int SMEM_BYTES = OFF_BAR + (NBARS * 8 + 8 + 15) / 16 * 16;
  static constexpr int THREADS = 320;
  static_assert(SMEM_BYTES <= 232448, "smem budget");
};

template <class Epi, int C, int NT_ = 192>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
k_gemm_pair(const uint8_t* __restrict__ Aimg,   // fp16 tile image of A (tokens, C)
            const Epi epi,
            const uint8_t* __restrict__ Wimg,   // [N/192][C/64][192 x 128B]
            long long M, int num_m_tiles, int num_n_tiles, int expflags) {
  using Cfg = GPairCfg<C, NT_>;
  extern __shared__ __align__(1024) uint8_t smem_gp[];
  uint8_t* smem = smem_gp;
  uint8_t* a_s = smem;
  uint8_t* w_s = smem + Cfg::OFF_W;
  float* patches = reinterpret_cast<float*>(smem + Cfg::OFF_PATCH);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* a_full = bars;                      // [NKB]
  uint64_t* a_empty = a_full + Cfg::NKB;        // [NKB]
  uint64_t* a_peer = a_empty + Cfg::NKB;        // [NKB] (used in the even CTA)
  uint64_t* w_full = a_peer + Cfg::NKB;         // [S]
  uint64_t* w_empty = w_full + Cfg::S;          // [S]
  uint64_t* w_peer = w_empty + Cfg::S;          // [S]   (used in the even CTA)
  uint64_t* acc_full = w_peer + Cfg::S;         // [2]
  uint64_t* acc_empty = acc_full + 2;           // [2]   (even CTA, 16 arrivals)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x / 32, 0), lane = threadIdx.x % 32;   // provably warp-uniform
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int num_super = (num_m_tiles + 1) / 2;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    for (int k = 0; k < Cfg::NKB; ++k) { mbar_init(&a_full[k], 1); mbar_init(&a_empty[k], 1); mbar_init(&a_peer[k], 1); }
    for (int s = 0; s < Cfg::S; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); mbar_init(&w_peer[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 16); }
    mbar_fence_init();
  }
  if (warp == 9) tmem_alloc_pair<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's barriers are initialised before anyone arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 8) {
    // ===================== loader =====================
    int s = 0; uint32_t ph = 0, tph = 0;
    for (int sup = pair; sup < num_super; sup += npairs, tph ^= 1) {
      int mt = 2 * sup + (int)rank;
      if (mt > num_m_tiles - 1) mt = num_m_tiles - 1;   // odd tile count: the idle half recomputes the last tile, stores nothing
      for (int nt = 0; nt < num_n_tiles; ++nt) {
        for (int kb = 0; kb < Cfg::NKB; ++kb) {
          if (nt == 0) {
            mbar_wait(&a_empty[kb], tph ^ 1);
            if (lane == 0) {
              mbar_arrive_expect_tx(&a_full[kb], G2_A_BYTES);
              bulk_g2s(a_s + kb * G2_A_BYTES, Aimg + ((size_t)mt * Cfg::NKB + kb) * G2_A_BYTES, G2_A_BYTES, &a_full[kb]);
            }
            __syncwarp();
          }
          mbar_wait(&w_empty[s], ph ^ 1);
          if (lane == 0) {
            mbar_arrive_expect_tx(&w_full[s], Cfg::W_HALF);
            bulk_g2s(w_s + s * Cfg::W_HALF, Wimg + ((size_t)nt * Cfg::NKB + kb) * Cfg::W_FULL + rank * Cfg::W_HALF,
                     Cfg::W_HALF, &w_full[s]);
          }
          __syncwarp();
          if (++s == Cfg::S) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 9 && rank == 0) {
    // ===================== MMA issuer (even CTA) =====================
    constexpr uint32_t idesc = make_idesc_f16(256, Cfg::NT);
    int s = 0; uint32_t ph = 0, tph = 0, cnt = 0;
    const uint32_t a_addr = smem_u32(a_s), w_addr = smem_u32(w_s);
    for (int sup = pair; sup < num_super; sup += npairs, tph ^= 1) {
      for (int nt = 0; nt < num_n_tiles; ++nt, ++cnt) {
        const uint32_t buf = cnt & 1, use = cnt >> 1;
        mbar_wait(&acc_empty[buf], (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * Cfg::NT;
        const bool last_nt = nt == num_n_tiles - 1;
        for (int kb = 0; kb < Cfg::NKB; ++kb) {
          if (nt == 0) { mbar_wait(&a_full[kb], tph); mbar_wait(&a_peer[kb], tph); }
          mbar_wait(&w_full[s], ph);
          mbar_wait(&w_peer[s], ph);
          tc_fence_after();
          const uint64_t da = make_desc_sw128(a_addr + kb * G2_A_BYTES);   // +2 in the address field = one K=16 step
          const uint64_t db = make_desc_sw128(w_addr + s * Cfg::W_HALF);
          if (elect_one()) {
            tc_mma_f16_pair(d_tmem, da, db, idesc, kb != 0 ? 1u : 0u);
            tc_mma_f16_pair(d_tmem, da + 2, db + 2, idesc, 1u);
            tc_mma_f16_pair(d_tmem, da + 4, db + 4, idesc, 1u);
            tc_mma_f16_pair(d_tmem, da + 6, db + 6, idesc, 1u);
            tc_commit_pair(&w_empty[s]);
            if (last_nt) tc_commit_pair(&a_empty[kb]);                 // this A k-block is not needed again
            if (kb == Cfg::NKB - 1) tc_commit_pair(&acc_full[buf]);
          }
          __syncwarp();
          if (++s == Cfg::S) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 9) {
    // ===================== completion relay (odd CTA) =====================
    const uint32_t r_a = mapa_u32(smem_u32(a_peer), 0);
    const uint32_t r_w = mapa_u32(smem_u32(w_peer), 0);
    int s = 0; uint32_t ph = 0, tph = 0;
    for (int sup = pair; sup < num_super; sup += npairs, tph ^= 1) {
      for (int nt = 0; nt < num_n_tiles; ++nt) {
        for (int kb = 0; kb < Cfg::NKB; ++kb) {
          if (nt == 0) {
            mbar_wait(&a_full[kb], tph);
            if (lane == 0) mbar_arrive_cluster(r_a + kb * 8);
            __syncwarp();
          }
          mbar_wait(&w_full[s], ph);
          if (lane == 0) mbar_arrive_cluster(r_w + s * 8);
          __syncwarp();
          if (++s == Cfg::S) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue warps 0..7 =====================
    const int q = warp & 3, part = warp >> 2;
    EpiCtx ctx;
    ctx.M = M; ctx.lane = lane; ctx.part = part; ctx.nparts = 2;
    ctx.patch = patches + warp * G2_PATCH_FLOATS;
    ctx.patch_s = smem_u32(ctx.patch);
    ctx.svec_s = 0;
    const uint32_t r_acc_empty = mapa_u32(smem_u32(acc_empty), 0);
    uint32_t cnt = 0;
    for (int sup = pair; sup < num_super; sup += npairs) {
      ctx.row0 = (long long)(2 * sup + (int)rank) * 128 + q * 32;
      for (int nt = 0; nt < num_n_tiles; ++nt, ++cnt) {
        const uint32_t buf = cnt & 1, use = cnt >> 1;
        ctx.n0 = nt * Cfg::NT;
        if constexpr (Epi::kHasPre) epi.template pre<Cfg::NT>(ctx);   // work that can start before the accumulator is ready
        mbar_wait(&acc_full[buf], use & 1);
        tc_fence_after();
        AccTmem2 acc{tmem_base + ((uint32_t)(q * 32) << 16) + buf * Cfg::NT};
        if (!(expflags & 1)) epi.template run<Cfg::NT>(acc, ctx);   // bit0: timing experiment, no epilogue
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(r_acc_empty + buf * 8);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // nobody leaves (or frees TMEM) while the peer can still signal / be signalled
  tc_fence_after();
  if (warp == 9) {
    __syncwarp();
    tmem_dealloc_pair<512>(tmem_base);
  }
}

template <class Epi, int C, int NT_ = 192>
int launch_gemm_pair(const uint8_t* Aimg, const Epi& epi, const uint8_t* Wimg, long long M, int N, int num_sms,
                     cudaStream_t st) {
  using Cfg = GPairCfg<C, NT_>;
  auto kern = k_gemm_pair<Epi, C, NT_>;
  static std::atomic<uint64_t> configured{0};   // one bit per device: the attribute is per (function, device)
  if (int rc = smem_opt_in(configured, reinterpret_cast<const void*>(kern), Cfg::SMEM_BYTES)) return rc;
  const int tiles = (int)((M + 127) / 128);
  const int supers = (tiles + 1) / 2;
  const int pairs = supers < num_sms / 2 ? supers : num_sms / 2;
#ifdef SKY_EXPERIMENTS
  static const int expflags = getenv("SKY_QKV_EXP") ? atoi(getenv("SKY_QKV_EXP")) : 0;  // timing experiments only
#else
  constexpr int expflags = 0;
#endif
  kern<<<2 * pairs, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(Aimg, epi, Wimg, M, tiles, N / Cfg::NT, expflags);
  SKY_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace sky
