// Fused transformer MLP on CTA pairs (tcgen05 cta_group::2), one 256-token super-tile per pair:
//
//     x += LayerNorm( GELU(xh W1^T + b1) W2^T + b2 ) * gamma + beta ;   xh <- fp16 image of x
//
// Same data flow as mlp_fused.cuh (the hidden activation lives in TMEM -> registers -> shared
// memory only), but every MMA is M=256 across two SMs: each CTA holds its own 128 token rows
// (A tile, hidden chunk, accumulators) and only HALF of every weight item, so the weight
// traffic L2 -> SM, the TMA writes into shared memory and the B-operand reads of the tensor
// core are halved per SM.  At C=384 the single-CTA kernel is shared-memory-bandwidth bound
// (DESIGN.md §7); this is the fix.
//
// Per CTA, 320 threads: warps 0..7 epilogue (GELU per chunk, LayerNorm + residual per tile),
// 8 loader (A tile + the weight ring), 9 MMA issuer in the even CTA / completion relay in the odd
// CTA.  1-D bulk copies cannot signal an mbarrier of the peer CTA, so the odd CTA's warp 9 forwards
// "my operand landed" to the issuer with remote mbarrier arrives, in the issuer's own consumption
// order.  tcgen05.commit multicasts "operands consumed / accumulator ready" to both CTAs.
//
//   TMEM columns (per CTA): acc1[ACC1_BUFS] (HC=128 each) | acc2 (C)
//     C=192: 2*128 + 192 = 448      C=384: 128 + 384 = 512 (acc1 single buffered: the epilogue
//     releases it as soon as its TMEM loads have landed in registers, before the GELU math)
#pragma once
#include <cstdio>
#include <cstdlib>

#include "gemm2.cuh"

namespace sky {

template <int C>
struct Mlp2Cfg {
  static constexpr int NKB = C / 64;
  static constexpr int HC = 128;
  static constexpr int NCH = 4 * C / HC;
  static constexpr int HKB = HC / 64;
  static constexpr int ACC1_BUFS = (2 * HC + C <= 512) ? 2 : 1;
  static constexpr int W1_FULL = HC * 128;       // one (chunk, k-block) item of the W1 image
  static constexpr int W1_HALF = W1_FULL / 2;    // this CTA's 64 rows of it
  // GEMM2 runs as NP MMAs of N2 output columns per K=16 step.  C=384 uses N2=128 so that a W2 item is
  // 8 KB like a W1 item and both share ONE ring in consumption order: with separate rings the idle
  // ring's slots hold nothing while the other one starves (measured: the issuer waited 2k cycles per
  // chunk on W2 with 2 slots, profiles/r1_mlp_pair.md).
  // (tried in round 2: N2 = 192 at C=384, i.e. two n-parts and 12 KB ring slots -> only 4 slots fit: mlp 5.74 vs 5.20 ms/step)
  static constexpr int N2 = C == 384 ? 128 : 192;
  static constexpr int NP = C / N2;
  static constexpr int W2_HALF = N2 / 2 * 128;   // this CTA's N2/2 rows of a (k-block, n-part) item of the W2 image
  static constexpr int SLOT = W1_HALF > W2_HALF ? W1_HALF : W2_HALF;
  static constexpr int S = C == 192 ? 8 : 7;
  static constexpr int A_BYTES = NKB * G2_A_BYTES;
  static constexpr int HID_BYTES = HKB * G2_A_BYTES;
  static constexpr int OFF_W = A_BYTES;
  static constexpr int OFF_HID = OFF_W + S * SLOT;
  static constexpr int PATCH_BYTES = 8 * G2_PATCH_FLOATS * 4;   // LN patches alias the hidden buffers
  static constexpr int HID_REGION = 2 * HID_BYTES;
  static constexpr int OFF_VEC = OFF_HID + HID_REGION;          // b1[4C]
  static constexpr int OFF_LNV = OFF_VEC + 4 * C * 4;           // bias2 | gamma | beta, C floats apart
  static constexpr int OFF_BAR = OFF_LNV + 3 * C * 4;
  static constexpr int NBARS = 3 + 3 * S + 8 + 2;
  static constexpr int SMEM_BYTES = OFF_BAR + (NBARS * 8 + 8 + 15) / 16 * 16;
  static constexpr int ACC2_COL = ACC1_BUFS * HC;
  static constexpr int THREADS = 320;
  static_assert(ACC1_BUFS * HC + C <= 512, "TMEM budget");
  static_assert(HID_REGION >= PATCH_BYTES, "LN patches must fit in the hidden buffers");
  static_assert(SLOT % 1024 == 0, "operand slots must keep the 1024-byte swizzle atom alignment");
  static_assert(SMEM_BYTES <= 232448, "smem budget");
};

template <int C, class EpiT = EpiLnRes>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
k_mlp_fused_pair(const uint8_t* __restrict__ xh_in,  // A image of x (tokens, C)
                 const EpiT epi,   // x (fp32), xh out image, b2, gamma, beta
                 const uint8_t* __restrict__ W1img,  // [4C/HC][C/64][HC x 128B]
                 const uint8_t* __restrict__ W2img,  // [1][4C/64][C x 128B]
                 const float* __restrict__ b1, long long M, int num_m_tiles, long long* dbg_in) {
  using Cfg = Mlp2Cfg<C>;
#ifdef SKY_EXPERIMENTS
  long long* const dbg = dbg_in;      // cycle accounting of the role warps (dev build, SKY_MLP_DBG)
#else
  constexpr long long* dbg = nullptr;  // product build: every timer below folds away
#endif
  long long tacc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define SKY_T2(i, stmt) do { long long _t0 = dbg ? clock64() : 0; stmt; if (dbg) tacc[i] += clock64() - _t0; } while (0)
  const long long t_begin = dbg ? clock64() : 0;
  extern __shared__ __align__(1024) uint8_t smem_pair[];
  uint8_t* smem = smem_pair;
  uint8_t* a_s = smem;
  uint8_t* w_s = smem + Cfg::OFF_W;
  uint8_t* hid_s = smem + Cfg::OFF_HID;
  float* b1s = reinterpret_cast<float*>(smem + Cfg::OFF_VEC);
  float* lnv = reinterpret_cast<float*>(smem + Cfg::OFF_LNV);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* a_full = bars + 0;
  uint64_t* a_empty = bars + 1;
  uint64_t* a_peer = bars + 2;                  //       (used in the even CTA)
  uint64_t* w_full = bars + 3;                  // [S]
  uint64_t* w_empty = w_full + Cfg::S;          // [S]
  uint64_t* w_peer = w_empty + Cfg::S;          // [S]   (used in the even CTA)
  uint64_t* acc1_full = w_peer + Cfg::S;        // [2]
  uint64_t* acc1_empty = acc1_full + 2;         // [2]   (even CTA, 16 arrivals)
  uint64_t* hid_full = acc1_empty + 2;          // [2]   (even CTA, 16 arrivals)
  uint64_t* hid_empty = hid_full + 2;           // [2]
  uint64_t* acc2_full = hid_empty + 2;
  uint64_t* acc2_empty = acc2_full + 1;         //       (even CTA, 16 arrivals)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc2_empty + 1);

  // the shuffle makes the warp index provably warp-uniform for the compiler: the role branches then are uniform control
  // flow and descriptor / address arithmetic of the issuer runs on the uniform datapath (CUTLASS canonical_warp_idx_sync)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x / 32, 0), lane = threadIdx.x % 32;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int num_super = (num_m_tiles + 1) / 2;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    mbar_init(a_full, 1); mbar_init(a_empty, 1); mbar_init(a_peer, 1);
    for (int s = 0; s < Cfg::S; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); mbar_init(&w_peer[s], 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc1_full[b], 1); mbar_init(&acc1_empty[b], 16);
      mbar_init(&hid_full[b], 16); mbar_init(&hid_empty[b], 1);
    }
    mbar_init(acc2_full, 1); mbar_init(acc2_empty, 16);
    mbar_fence_init();
  }
  if (warp == 9) tmem_alloc_pair<512>(tmem_ptr);
  for (int i = threadIdx.x; i < 4 * C; i += blockDim.x) b1s[i] = b1[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) { lnv[i] = epi.bias[i]; lnv[C + i] = epi.gamma[i]; lnv[2 * C + i] = epi.beta[i]; }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's barriers are initialised before anyone arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // Weight items travel through one ring in the order the issuer consumes them:
  //   per tile:  W1(0) | W1(1) W2(0) | W1(2) W2(1) | ... | W1(NCH-1) W2(NCH-2) | W2(NCH-1)
  // W1(j) = NKB items (k-blocks of chunk j), W2(j) = HKB*NP items (k-block, n-part).  Waits are CTA-scope
  // acquires: a cluster-scope acquire adds CCTL.IVALL after every wait and no thread here reads
  // peer-written memory through the generic proxy.
  if (warp == 8) {
    // ===================== loader: own A tile + own half of every weight item =====================
    int s = 0; uint32_t ph = 0; uint32_t tph = 0;
    auto put = [&](const uint8_t* src, uint32_t bytes) {
      mbar_wait(&w_empty[s], ph ^ 1);
      if (lane == 0) {
        mbar_arrive_expect_tx(&w_full[s], bytes);
        bulk_g2s(w_s + s * Cfg::SLOT, src, bytes, &w_full[s]);
      }
      __syncwarp();
      if (++s == Cfg::S) { s = 0; ph ^= 1; }
    };
    auto put_w2 = [&](int j) {
      for (int kb2 = 0; kb2 < Cfg::HKB; ++kb2)
        for (int np = 0; np < Cfg::NP; ++np)
          put(W2img + ((size_t)(j * Cfg::HKB + kb2) * C + np * Cfg::N2 + rank * (Cfg::N2 / 2)) * 128, Cfg::W2_HALF);
    };
    for (int sup = pair; sup < num_super; sup += npairs, tph ^= 1) {
      int mt = 2 * sup + (int)rank;
      if (mt > num_m_tiles - 1) mt = num_m_tiles - 1;   // odd tile count: the idle half recomputes the last tile, stores nothing
      mbar_wait(a_empty, tph ^ 1);
      if (lane == 0) {
        mbar_arrive_expect_tx(a_full, Cfg::A_BYTES);
        for (int kb = 0; kb < Cfg::NKB; ++kb)
          bulk_g2s(a_s + kb * G2_A_BYTES, xh_in + ((size_t)mt * Cfg::NKB + kb) * G2_A_BYTES, G2_A_BYTES, a_full);
      }
      __syncwarp();
      for (int j = 0; j < Cfg::NCH; ++j) {
        for (int kb = 0; kb < Cfg::NKB; ++kb)
          put(W1img + ((size_t)j * Cfg::NKB + kb) * Cfg::W1_FULL + rank * Cfg::W1_HALF, Cfg::W1_HALF);
        if (j >= 1) put_w2(j - 1);
      }
      put_w2(Cfg::NCH - 1);
    }
  } else if (warp == 9 && rank == 0) {
    // ===================== MMA issuer (even CTA) =====================
    constexpr uint32_t idesc1 = make_idesc_f16(256, Cfg::HC);
    constexpr uint32_t idesc2 = make_idesc_f16(256, Cfg::N2);
    int s = 0; uint32_t ph = 0;
    uint32_t tph = 0;        // tile parity (a_full, acc2)
    uint32_t cnt = 0;        // global chunk counter
    const uint32_t a_addr = smem_u32(a_s);
    const uint32_t hid_addr = smem_u32(hid_s);
    const uint32_t w_addr = smem_u32(w_s);
    // one weight item = 4 MMAs (K = 64); descriptors are computed by the whole warp (uniform datapath);
    // +2 in the descriptor's address field = +32 bytes = one K=16 step
    auto item = [&](uint32_t a_bytes, uint32_t d_tmem, uint32_t idesc, uint32_t acc0, uint64_t* extra0, uint64_t* extra1) {
      SKY_T2(2, mbar_wait(&w_full[s], ph));
      SKY_T2(3, mbar_wait(&w_peer[s], ph));
      tc_fence_after();
      const uint64_t da = make_desc_sw128(a_bytes);
      const uint64_t db = make_desc_sw128(w_addr + s * Cfg::SLOT);
      if (elect_one()) {
        tc_mma_f16_pair(d_tmem, da, db, idesc, acc0);
        tc_mma_f16_pair(d_tmem, da + 2, db + 2, idesc, 1u);
        tc_mma_f16_pair(d_tmem, da + 4, db + 4, idesc, 1u);
        tc_mma_f16_pair(d_tmem, da + 6, db + 6, idesc, 1u);
        tc_commit_pair(&w_empty[s]);
        if (extra0) tc_commit_pair(extra0);
        if (extra1) tc_commit_pair(extra1);
      }
      __syncwarp();
      if (++s == Cfg::S) { s = 0; ph ^= 1; }
    };
    auto gemm2 = [&](uint32_t ci /*global chunk id*/, bool first_of_tile, bool last_of_tile) {
      const uint32_t hb = ci & 1, use = ci >> 1;
      SKY_T2(4, mbar_wait(&hid_full[hb], use & 1));
      if (first_of_tile) SKY_T2(7, mbar_wait(acc2_empty, tph ^ 1));
      tc_fence_after();
      for (int kb2 = 0; kb2 < Cfg::HKB; ++kb2)
        for (int np = 0; np < Cfg::NP; ++np) {
          const bool last = kb2 == Cfg::HKB - 1 && np == Cfg::NP - 1;
          item(hid_addr + hb * Cfg::HID_BYTES + kb2 * G2_A_BYTES, tmem_base + Cfg::ACC2_COL + np * Cfg::N2, idesc2,
               (!first_of_tile || kb2 > 0) ? 1u : 0u, last ? &hid_empty[hb] : nullptr, (last && last_of_tile) ? acc2_full : nullptr);
        }
    };
    for (int sup = pair; sup < num_super; sup += npairs, tph ^= 1) {
      SKY_T2(0, mbar_wait(a_full, tph));
      SKY_T2(0, mbar_wait(a_peer, tph));
      tc_fence_after();
      for (int j = 0; j < Cfg::NCH; ++j, ++cnt) {
        const uint32_t ab = cnt % Cfg::ACC1_BUFS, use = cnt / Cfg::ACC1_BUFS;
        SKY_T2(1, mbar_wait(&acc1_empty[ab], (use & 1) ^ 1));
        tc_fence_after();
        for (int kb = 0; kb < Cfg::NKB; ++kb) {
          const bool last = kb == Cfg::NKB - 1;
          item(a_addr + kb * G2_A_BYTES, tmem_base + ab * Cfg::HC, idesc1, kb != 0 ? 1u : 0u, last ? &acc1_full[ab] : nullptr,
               (last && j == Cfg::NCH - 1) ? a_empty : nullptr);
        }
        if (j >= 1) gemm2(cnt - 1, j == 1, false);
      }
      gemm2(cnt - 1, Cfg::NCH == 1, true);
    }
  } else if (warp == 9) {
    // ===================== completion relay (odd CTA) =====================
    // 1-D bulk copies cannot signal the peer's mbarrier: forward "my half of the operand is in shared memory" to the issuer
    const uint32_t r_a = mapa_u32(smem_u32(a_peer), 0);
    const uint32_t r_w = mapa_u32(smem_u32(w_peer), 0);
    int s = 0; uint32_t ph = 0, tph = 0;
    constexpr int ITEMS_PER_TILE = Cfg::NCH * (Cfg::NKB + Cfg::HKB * Cfg::NP);
    for (int sup = pair; sup < num_super; sup += npairs, tph ^= 1) {
      mbar_wait(a_full, tph);
      if (lane == 0) mbar_arrive_cluster(r_a);
      __syncwarp();
      for (int i = 0; i < ITEMS_PER_TILE; ++i) {
        mbar_wait(&w_full[s], ph);
        if (lane == 0) mbar_arrive_cluster(r_w + s * 8);
        __syncwarp();
        if (++s == Cfg::S) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps 0..7 =====================
    const int q = warp & 3, part = warp >> 2;
    EpiCtx ctx;
    ctx.M = M; ctx.lane = lane; ctx.part = part; ctx.nparts = 2; ctx.n0 = 0;
    ctx.patch = reinterpret_cast<float*>(hid_s) + warp * G2_PATCH_FLOATS;   // 8 x 4.2 KB inside the 64 KB hidden region
    ctx.patch_s = smem_u32(ctx.patch);
    ctx.svec_s = smem_u32(lnv);
    ctx.vstride = C;
    const uint32_t b1s_s = smem_u32(b1s);
    const uint32_t r_acc1_empty = mapa_u32(smem_u32(acc1_empty), 0);
    const uint32_t r_hid_full = mapa_u32(smem_u32(hid_full), 0);
    const uint32_t r_acc2_empty = mapa_u32(smem_u32(acc2_empty), 0);
    uint32_t cnt = 0, tph = 0;
    const uint32_t r = q * 32 + lane;
    for (int sup = pair; sup < num_super; sup += npairs, tph ^= 1) {
      const int mt = 2 * sup + (int)rank;
      ctx.row0 = (long long)mt * 128 + q * 32;
      epi.template prefetch<C>(ctx);  // residual rows -> L2 while the tile's GEMMs run
      for (int j = 0; j < Cfg::NCH; ++j, ++cnt) {
        const uint32_t ab = cnt % Cfg::ACC1_BUFS, ause = cnt / Cfg::ACC1_BUFS;
        const uint32_t hb = cnt & 1, huse = cnt >> 1;
        SKY_T2(0, mbar_wait(&acc1_full[ab], ause & 1));
        tc_fence_after();
        const long long _tg = dbg ? clock64() : 0;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + ab * Cfg::HC + part * 64;
        float v0[32], v1[32];
        __syncwarp();
        tmem_ld32_nowait(taddr, v0);
        tmem_ld32_nowait(taddr + 32, v1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(r_acc1_empty + ab * 8);   // accumulator is in registers: GEMM1 of the next chunk may start
        SKY_T2(1, mbar_wait(&hid_empty[hb], (huse & 1) ^ 1));
        uint8_t* kbase = hid_s + hb * Cfg::HID_BYTES + part * G2_A_BYTES;   // this warp's 64 columns = k-block `part` of the chunk
        const uint32_t bb = b1s_s + (j * Cfg::HC + part * 64) * 4;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b4 = lds_f32x4_ro(bb + i * 4);
          gelu_erf_x2(v0[i], v0[i + 1], b4.x, b4.y);
          gelu_erf_x2(v0[i + 2], v0[i + 3], b4.z, b4.w);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 pk;
          pk.x = pack_half2(v0[8 * i], v0[8 * i + 1]); pk.y = pack_half2(v0[8 * i + 2], v0[8 * i + 3]);
          pk.z = pack_half2(v0[8 * i + 4], v0[8 * i + 5]); pk.w = pack_half2(v0[8 * i + 6], v0[8 * i + 7]);
          *reinterpret_cast<uint4*>(kbase + sw128_offset(r, i)) = pk;
        }
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b4 = lds_f32x4_ro(bb + (32 + i) * 4);
          gelu_erf_x2(v1[i], v1[i + 1], b4.x, b4.y);
          gelu_erf_x2(v1[i + 2], v1[i + 3], b4.z, b4.w);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 pk;
          pk.x = pack_half2(v1[8 * i], v1[8 * i + 1]); pk.y = pack_half2(v1[8 * i + 2], v1[8 * i + 3]);
          pk.z = pack_half2(v1[8 * i + 4], v1[8 * i + 5]); pk.w = pack_half2(v1[8 * i + 6], v1[8 * i + 7]);
          *reinterpret_cast<uint4*>(kbase + sw128_offset(r, 4 + i)) = pk;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(r_hid_full + hb * 8);
        if (dbg) tacc[2] += clock64() - _tg;
      }
      // ---- LayerNorm + residual on the finished acc2 tile ----
      SKY_T2(3, mbar_wait(acc2_full, tph));
      tc_fence_after();
      const long long _tl = dbg ? clock64() : 0;
      AccTmem2 acc{tmem_base + ((uint32_t)(q * 32) << 16) + Cfg::ACC2_COL};
      epi.template run<C>(acc, ctx);
      tc_fence_before();
      // the LN patches alias the hidden buffers: no epilogue warp may start the next tile's
      // GELU stores before every warp has left its patch
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (lane == 0) mbar_arrive_cluster(r_acc2_empty);
      if (dbg) tacc[4] += clock64() - _tl;
    }
  }
  if (dbg && lane == 0 && (warp == 0 || warp == 9) && blockIdx.x < 4) {
    tacc[15] = clock64() - t_begin;
    for (int i = 0; i < 16; ++i) dbg[(blockIdx.x * 2 + (warp == 9)) * 16 + i] = tacc[i];
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

template <int C, class EpiT = EpiLnRes>
int launch_mlp_fused_pair(const uint8_t* xh_in, const EpiT& epi, const uint8_t* W1img,
                          const uint8_t* W2img, const float* b1, long long M, int num_sms, cudaStream_t st) {
  static long long* dbg = nullptr;
#ifdef SKY_EXPERIMENTS
  static int dbg_runs = 0;
  if (getenv("SKY_MLP_DBG") && !dbg) cudaMallocManaged(&dbg, 256 * 8);
#endif
  using Cfg = Mlp2Cfg<C>;
  auto kern = k_mlp_fused_pair<C, EpiT>;
  static std::atomic<uint64_t> configured{0};   // one bit per device: the attribute is per (function, device)
  if (int rc = smem_opt_in(configured, reinterpret_cast<const void*>(kern), Cfg::SMEM_BYTES)) return rc;
  const int tiles = (int)((M + 127) / 128);
  const int supers = (tiles + 1) / 2;
  const int pairs = supers < num_sms / 2 ? supers : num_sms / 2;
  kern<<<2 * pairs, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(xh_in, epi, W1img, W2img, b1, M, tiles, dbg);
#ifdef SKY_EXPERIMENTS
  if (dbg && dbg_runs < 2) {
    ++dbg_runs;
    cudaDeviceSynchronize();
    for (int b = 0; b < 2; ++b) {   // CTA 0 = issuer, CTA 1 = its peer
      const long long* e = dbg + (b * 2) * 16; const long long* m = dbg + (b * 2 + 1) * 16;
      printf("[mlp-pair C=%d cta %d, %d supertiles/pair, total %lld] EPI wait_acc1 %lld wait_hidempty %lld gelu(incl) %lld wait_acc2 %lld ln %lld | "
             "W9 wait_a %lld acc1empty %lld w %lld wpeer %lld hidfull %lld acc2empty %lld\n", C, b,
             (supers + pairs - 1) / pairs, e[15], e[0], e[1], e[2], e[3], e[4], m[0], m[1], m[2], m[3], m[4], m[7]);
    }
  }
#endif
  SKY_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace sky
