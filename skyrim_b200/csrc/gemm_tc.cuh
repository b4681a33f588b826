// Persistent warp-specialised tcgen05 GEMM for sm_100a:
//
//      D[M, N] = A(prod)[M, K] * W[N, K]^T  ->  epilogue(epi)
//
//   * A operand: gathered / normalised / converted fp32->fp16 by four producer warps
//     straight into a SWIZZLE_128B K-major smem tile (no intermediate HBM copy of the
//     windowed / im2col'd activations ever exists);
//   * B operand: weights pre-packed at load time as the exact smem tile image
//     (prep.cu::k_pack_weight_image), fetched with one 1-D bulk copy (TMA engine,
//     cp.async.bulk + mbarrier complete_tx) per k-block;
//   * one elected thread issues tcgen05.mma (kind::f16, M=128, fp32 accumulators in TMEM),
//     tcgen05.commit frees smem stages and publishes the accumulator;
//   * four epilogue warps read their TMEM lane quarter (tcgen05.ld 32x32b) and run the fused
//     epilogue; accumulators are double buffered in TMEM when 2*BLOCK_N <= 512 columns so
//     the epilogue of tile i overlaps the MMAs of tile i+1.
#pragma once
#include "pangu_ops.cuh"

namespace sky {

constexpr int TC_BLOCK_M = 128;
constexpr int TC_BLOCK_K = 64;                      // halves: one 128-byte swizzle row
constexpr int TC_A_BYTES = TC_BLOCK_M * 128;        // 16 KB
constexpr int TC_THREADS = 320;                     // 4 epilogue + 4 producer + MMA + B-loader warps
constexpr int TC_NUM_PRODUCER_THREADS = 128;

template <int BLOCK_N>
struct TcCfg {
  static constexpr int N_INST = BLOCK_N <= 256 ? BLOCK_N : BLOCK_N / 2;  // UMMA N per instruction
  static constexpr int N_SPLIT = BLOCK_N / N_INST;
  static constexpr int B_BYTES = BLOCK_N * 128;
  static constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES >= 6 ? 6 : (200 * 1024) / STAGE_BYTES;
  static constexpr int NBUF = 2 * BLOCK_N <= 512 ? 2 : 1;
  static constexpr int TMEM_COLS = NBUF * BLOCK_N <= 32 ? 32 : NBUF * BLOCK_N <= 64 ? 64
                                   : NBUF * BLOCK_N <= 128 ? 128 : NBUF * BLOCK_N <= 256 ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(N_INST % 16 == 0 && N_INST <= 256, "UMMA shape");
  static_assert(STAGE_BYTES % 1024 == 0, "stage alignment");
};

struct AccTmem {
  uint32_t taddr;  // lane quarter base | column base
  __device__ void load32(int c, float (&v)[32]) const { tmem_ld32(taddr + (uint32_t)c, v); }
};

template <class Prod, class Epi, int BLOCK_N>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_gemm_tc(const Prod prod, const Epi epi, const uint8_t* __restrict__ Wimg, long long M,
          int num_kb, int num_m_tiles, int num_n_tiles) {
  using Cfg = TcCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;                       // [STAGES]  producers(128) + B loader(1 + tx)
  uint64_t* empty = bars + Cfg::STAGES;        // [STAGES]  tcgen05.commit
  uint64_t* tmem_full = bars + 2 * Cfg::STAGES;       // [2]
  uint64_t* tmem_empty = bars + 2 * Cfg::STAGES + 2;  // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::STAGES + 4);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x / 32, 0), lane = threadIdx.x % 32;  // provably warp-uniform
  const int num_tiles = num_m_tiles * num_n_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full[s], TC_NUM_PRODUCER_THREADS + 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 128);
    }
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp >= 4 && warp < 8) {
    // ============================ A producers ============================
    // thread = tile row, the 8 chunks of a k-block in sequence: for one chunk the 32 lanes of a warp are 32
    // consecutive tokens, i.e. one contiguous 512-byte run of the source field per load instruction, and the
    // 16-byte shared-memory stores of a quarter warp fall into 8 different swizzled chunk slots
    const int row = threadIdx.x - 128;
    int s = 0; uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const long long m0 = (long long)(tile / num_n_tiles) * TC_BLOCK_M;
      const RowInfo ri = prod.prep(m0 + row);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* a_s = smem + s * Cfg::STAGE_BYTES;
        uint4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = prod.load8(ri, kb * TC_BLOCK_K + i * 8);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          *reinterpret_cast<uint4*>(a_s + sw128_offset(row, i)) = v[i];
        fence_proxy_async_smem();
        mbar_arrive(&full[s]);
        if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 9) {
    // ============================ B loader (TMA engine) ============================
    int s = 0; uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int nt = tile % num_n_tiles;
      const uint8_t* src = Wimg + (size_t)nt * num_kb * Cfg::B_BYTES;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty[s], ph ^ 1);
        if (lane == 0) {
          mbar_arrive_expect_tx(&full[s], Cfg::B_BYTES);
          bulk_g2s(smem + s * Cfg::STAGE_BYTES + TC_A_BYTES, src + (size_t)kb * Cfg::B_BYTES, Cfg::B_BYTES,
                   &full[s]);
        }
        __syncwarp();
        if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 8) {
    // ============================ MMA issuer ============================
    constexpr uint32_t idesc = make_idesc_f16(TC_BLOCK_M, Cfg::N_INST);
    int s = 0; uint32_t ph = 0; int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int buf = it % Cfg::NBUF;
      const uint32_t use = (uint32_t)(it / Cfg::NBUF);
      mbar_wait(&tmem_empty[buf], (use & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BLOCK_N);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        {  // descriptors on the uniform datapath, one elected thread issues the MMAs and the commits that track them
          const uint32_t a_addr = smem_u32(smem + s * Cfg::STAGE_BYTES);
          const uint64_t da = make_desc_sw128(a_addr);
          const uint64_t db = make_desc_sw128(a_addr + TC_A_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < TC_BLOCK_K / 16; ++k) {
#pragma unroll
              for (int ni = 0; ni < Cfg::N_SPLIT; ++ni)
                tc_mma_f16(d_tmem + ni * Cfg::N_INST, da + 2 * k, db + (uint64_t)(ni * Cfg::N_INST * 8 + 2 * k), idesc,
                           (kb | k) != 0 ? 1u : 0u);
            }
            tc_commit(&empty[s]);
            if (kb == num_kb - 1) tc_commit(&tmem_full[buf]);
          }
        }
        __syncwarp();
        if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ============================ epilogue warps 0..3 ============================
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int buf = it % Cfg::NBUF;
      const uint32_t use = (uint32_t)(it / Cfg::NBUF);
      const long long m0 = (long long)(tile / num_n_tiles) * TC_BLOCK_M;
      const int n0 = (tile % num_n_tiles) * BLOCK_N;
      mbar_wait(&tmem_full[buf], use & 1);
      tc_fence_after();
      AccTmem acc{tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * BLOCK_N)};
      epi.run(acc, m0 + warp * 32 + lane, n0, BLOCK_N);
      tc_fence_before();
      mbar_arrive(&tmem_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 8) {
    __syncwarp();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <class Prod, class Epi, int BLOCK_N>
int launch_gemm_tc(const Prod& prod, const Epi& epi, const uint8_t* Wimg, long long M, int N, int Kp,
                   int num_sms, cudaStream_t st) {
  using Cfg = TcCfg<BLOCK_N>;
  auto kern = k_gemm_tc<Prod, Epi, BLOCK_N>;
  static std::atomic<uint64_t> configured{0};   // one bit per device: the attribute is per (function, device)
  if (int rc = smem_opt_in(configured, reinterpret_cast<const void*>(kern), Cfg::SMEM_BYTES)) return rc;
  const int num_m_tiles = (int)((M + TC_BLOCK_M - 1) / TC_BLOCK_M);
  const int num_n_tiles = N / BLOCK_N;
  const int tiles = num_m_tiles * num_n_tiles;
  const int grid = tiles < num_sms ? tiles : num_sms;
  kern<<<grid, TC_THREADS, Cfg::SMEM_BYTES, st>>>(prod, epi, Wimg, M, Kp / TC_BLOCK_K, num_m_tiles, num_n_tiles);
  SKY_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace sky
