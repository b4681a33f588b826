// Fused transformer MLP for one 128-token tile (sm_100a):
//
//     x += LayerNorm( GELU(xh W1^T + b1) W2^T + b2 ) * gamma + beta ;   xh <- fp16 image of x
//
// The 4C-wide hidden activation never leaves the SM: it is produced HC columns at a time as a
// TMEM accumulator (GEMM1), pulled through registers for bias + GELU, written to shared memory
// as the SWIZZLE_128B A-operand of GEMM2 and consumed from there.  Unfused, the hidden tensor
// costs 2 x (tokens x 4C x 2 B) of HBM traffic per block and makes both GEMMs memory bound.
//
// Warp roles (352 threads): 0..7 epilogue (GELU per chunk, LayerNorm+residual per tile),
// 8 loader of A and the W1 ring, 9 MMA issuer, 10 loader of the W2 ring.  All operands arrive
// by 1-D bulk copies of pre-built tile images.
//
//   TMEM columns: acc1[0] | acc1[1] | acc2        (2*HC + C  <= 512)
//   HC = 64: C = 192 -> 12 chunks, C = 384 -> 24 chunks
#pragma once
#include <cstdio>
#include <cstdlib>

#include "gemm2.cuh"

namespace sky {

template <int C>
struct MlpCfg {
  static constexpr int NKB = C / 64;
  static constexpr int HC = C == 192 ? 128 : 64;  // N of GEMM1: an SS-mode MMA re-reads its 4 KB A slice per K=16 step, so small N is smem-bandwidth bound
  static constexpr int NCH = 4 * C / HC;
  static constexpr int HKB = HC / 64;
  static constexpr int NH = C / 192;
  static constexpr int W1_ITEM = HC * 128;
  static constexpr int W2_ITEM = 192 * 128;
  static constexpr int S1 = C == 192 ? 3 : 4;  // ring depth is what hides the ~2 us L2 latency of a weight item
  static constexpr int S2 = 2;
  static constexpr int A_BYTES = NKB * G2_A_BYTES;
  static constexpr int HID_BYTES = HKB * G2_A_BYTES;
  static constexpr int OFF_W1 = A_BYTES;
  static constexpr int OFF_W2 = OFF_W1 + S1 * W1_ITEM;
  static constexpr int OFF_HID = OFF_W2 + S2 * W2_ITEM;
  static constexpr int PATCH_BYTES = 8 * G2_PATCH_FLOATS * 4;   // LN patches alias the hidden buffers
  static constexpr int HID_REGION = ((2 * HID_BYTES > PATCH_BYTES ? 2 * HID_BYTES : PATCH_BYTES) + 1023) / 1024 * 1024;
  static constexpr int OFF_VEC = OFF_HID + HID_REGION;          // b1[4C] b2[C] floats
  static constexpr int OFF_LNV = OFF_VEC + 4 * C * 4;           // bias2 | gamma | beta, 512 floats apart
  static constexpr int OFF_BAR = OFF_LNV + 3 * 512 * 4;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  static constexpr int ACC2_COL = 2 * HC;
  static constexpr int THREADS = 352;
  static_assert(2 * HC + C <= 512, "TMEM budget");
  static_assert(SMEM_BYTES <= 232448, "smem budget");
};

template <int C, class EpiT = EpiLnRes>
__global__ void __launch_bounds__(352, 1)
k_mlp_fused(const uint8_t* __restrict__ xh_in,  // A image of x (tokens, C)
            const EpiT epi,   // x (fp32), xh out image, b2, gamma, beta
            const uint8_t* __restrict__ W1img,  // [4C/HC][C/64][HC x 128B]
            const uint8_t* __restrict__ W2img,  // [1][4C/64][C x 128B]
            const float* __restrict__ b1, long long M, int num_m_tiles, long long* dbg, int expflags) {
  using Cfg = MlpCfg<C>;
  long long tacc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define SKY_T(i, stmt) do { long long _t0 = dbg ? clock64() : 0; stmt; if (dbg) tacc[i] += clock64() - _t0; } while (0)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_s = smem;
  uint8_t* w1_s = smem + Cfg::OFF_W1;
  uint8_t* w2_s = smem + Cfg::OFF_W2;
  uint8_t* hid_s = smem + Cfg::OFF_HID;
  float* b1s = reinterpret_cast<float*>(smem + Cfg::OFF_VEC);
  float* lnv = reinterpret_cast<float*>(smem + Cfg::OFF_LNV);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* a_full = bars + 0;
  uint64_t* a_empty = bars + 1;
  uint64_t* w1_full = bars + 2;                 // [S1]
  uint64_t* w1_empty = w1_full + Cfg::S1;       // [S1]
  uint64_t* w2_full = w1_empty + Cfg::S1;       // [S2]
  uint64_t* w2_empty = w2_full + Cfg::S2;       // [S2]
  uint64_t* acc1_full = w2_empty + Cfg::S2;     // [2]
  uint64_t* acc1_empty = acc1_full + 2;         // [2]
  uint64_t* hid_full = acc1_empty + 2;          // [2]
  uint64_t* hid_empty = hid_full + 2;           // [2]
  uint64_t* acc2_full = hid_empty + 2;
  uint64_t* acc2_empty = acc2_full + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc2_empty + 1);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;

  if (threadIdx.x == 0) {
    mbar_init(a_full, 1); mbar_init(a_empty, 1);
    for (int s = 0; s < Cfg::S1; ++s) { mbar_init(&w1_full[s], 1); mbar_init(&w1_empty[s], 1); }
    for (int s = 0; s < Cfg::S2; ++s) { mbar_init(&w2_full[s], 1); mbar_init(&w2_empty[s], 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc1_full[b], 1); mbar_init(&acc1_empty[b], 8);
      mbar_init(&hid_full[b], 8); mbar_init(&hid_empty[b], 1);
    }
    mbar_init(acc2_full, 1); mbar_init(acc2_empty, 8);
    mbar_fence_init();
  }
  if (warp == 9) tmem_alloc<512>(tmem_ptr);
  for (int i = threadIdx.x; i < 4 * C; i += blockDim.x) b1s[i] = b1[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) { lnv[i] = epi.bias[i]; lnv[512 + i] = epi.gamma[i]; lnv[1024 + i] = epi.beta[i]; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 8) {
    // ===================== loader: A tile + W1 ring =====================
    int s = 0; uint32_t ph = 0; uint32_t tph = 0;
    for (int mt = blockIdx.x; mt < num_m_tiles; mt += gridDim.x, tph ^= 1) {
      mbar_wait(a_empty, tph ^ 1);
      if (lane == 0) {
        mbar_arrive_expect_tx(a_full, Cfg::A_BYTES);
        for (int kb = 0; kb < Cfg::NKB; ++kb)
          bulk_g2s(a_s + kb * G2_A_BYTES, xh_in + ((size_t)mt * Cfg::NKB + kb) * G2_A_BYTES, G2_A_BYTES, a_full);
      }
      __syncwarp();
      for (int j = 0; j < Cfg::NCH; ++j) {
        for (int kb = 0; kb < Cfg::NKB; ++kb) {
          mbar_wait(&w1_empty[s], ph ^ 1);
          if (lane == 0) {
            mbar_arrive_expect_tx(&w1_full[s], Cfg::W1_ITEM);
            bulk_g2s(w1_s + s * Cfg::W1_ITEM, W1img + ((size_t)j * Cfg::NKB + kb) * Cfg::W1_ITEM, Cfg::W1_ITEM,
                     &w1_full[s]);
          }
          __syncwarp();
          if (++s == Cfg::S1) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 10) {
    // ===================== loader: W2 ring =====================
    int s = 0; uint32_t ph = 0;
    for (int mt = blockIdx.x; mt < num_m_tiles; mt += gridDim.x) {
      for (int j = 0; j < Cfg::NCH; ++j) {
        for (int kb2 = 0; kb2 < Cfg::HKB; ++kb2) {
          for (int nh = 0; nh < Cfg::NH; ++nh) {
            mbar_wait(&w2_empty[s], ph ^ 1);
            if (lane == 0) {
              mbar_arrive_expect_tx(&w2_full[s], Cfg::W2_ITEM);
              bulk_g2s(w2_s + s * Cfg::W2_ITEM,
                       W2img + ((size_t)(j * Cfg::HKB + kb2) * C + nh * 192) * 128, Cfg::W2_ITEM, &w2_full[s]);
            }
            __syncwarp();
            if (++s == Cfg::S2) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc1 = make_idesc_f16(128, Cfg::HC);
    constexpr uint32_t idesc2 = make_idesc_f16(128, 192);
    int s1 = 0; uint32_t ph1 = 0; int s2 = 0; uint32_t ph2 = 0;
    uint32_t tph = 0;        // tile parity (a_full, acc2)
    uint32_t cnt = 0;        // global chunk counter (acc1 / hid buffer parities)
    const uint32_t a_addr = smem_u32(a_s);
    const uint32_t hid_addr = smem_u32(hid_s);
    auto gemm2 = [&](uint32_t ci /*global chunk id*/, bool first_of_tile) {
      const uint32_t buf = ci & 1, use = ci >> 1;
      SKY_T(3, mbar_wait(&hid_full[buf], use & 1));
      if (first_of_tile) mbar_wait(acc2_empty, tph ^ 1);
      tc_fence_after();
      for (int kb2 = 0; kb2 < Cfg::HKB; ++kb2) {
        for (int nh = 0; nh < Cfg::NH; ++nh) {
          SKY_T(4, mbar_wait(&w2_full[s2], ph2));
          tc_fence_after();
          if (lane == 0) {
            const uint32_t b_addr = smem_u32(w2_s + s2 * Cfg::W2_ITEM);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t da = make_desc_sw128(hid_addr + buf * Cfg::HID_BYTES + kb2 * G2_A_BYTES + k * 32);
              const uint64_t db = make_desc_sw128(b_addr + k * 32);
              tc_mma_f16(tmem_base + Cfg::ACC2_COL + nh * 192, da, db, idesc2,
                         (!first_of_tile || kb2 > 0 || k > 0) ? 1u : 0u);
            }
            tc_commit(&w2_empty[s2]);
          }
          __syncwarp();
          if (++s2 == Cfg::S2) { s2 = 0; ph2 ^= 1; }
        }
      }
      if (lane == 0) tc_commit(&hid_empty[buf]);
      __syncwarp();
    };
    for (int mt = blockIdx.x; mt < num_m_tiles; mt += gridDim.x, tph ^= 1) {
      SKY_T(0, mbar_wait(a_full, tph));
      tc_fence_after();
      for (int j = 0; j < Cfg::NCH; ++j, ++cnt) {
        const uint32_t buf = cnt & 1, use = cnt >> 1;
        SKY_T(1, mbar_wait(&acc1_empty[buf], (use & 1) ^ 1));
        tc_fence_after();
        for (int kb = 0; kb < Cfg::NKB; ++kb) {
          SKY_T(2, mbar_wait(&w1_full[s1], ph1));
          tc_fence_after();
          if (lane == 0) {
            const uint32_t b_addr = smem_u32(w1_s + s1 * Cfg::W1_ITEM);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t da = make_desc_sw128(a_addr + kb * G2_A_BYTES + k * 32);
              const uint64_t db = make_desc_sw128(b_addr + k * 32);
              tc_mma_f16(tmem_base + buf * Cfg::HC, da, db, idesc1, (kb | k) != 0 ? 1u : 0u);
            }
            tc_commit(&w1_empty[s1]);
            if (kb == Cfg::NKB - 1) {
              tc_commit(&acc1_full[buf]);
              if (j == Cfg::NCH - 1) tc_commit(a_empty);
            }
          }
          __syncwarp();
          if (++s1 == Cfg::S1) { s1 = 0; ph1 ^= 1; }
        }
        if (j >= 1) gemm2(cnt - 1, j == 1);
      }
      gemm2(cnt - 1, Cfg::NCH == 1);
      if (lane == 0) tc_commit(acc2_full);
      __syncwarp();
    }
  } else {
    // ===================== epilogue warps 0..7 =====================
    const int q = warp & 3, part = warp >> 2;
    EpiCtx ctx;
    ctx.M = M; ctx.lane = lane; ctx.part = part; ctx.nparts = 2; ctx.n0 = 0;
    ctx.patch = reinterpret_cast<float*>(hid_s) + warp * G2_PATCH_FLOATS;
    ctx.patch_s = smem_u32(ctx.patch);
    ctx.svec_s = smem_u32(lnv);
    const uint32_t b1s_s = smem_u32(b1s);
    uint32_t cnt = 0, tph = 0;
    constexpr int COLS_PER_WARP = Cfg::HC / 2;
    for (int mt = blockIdx.x; mt < num_m_tiles; mt += gridDim.x, tph ^= 1) {
      ctx.row0 = (long long)mt * 128 + q * 32;
      epi.template prefetch<C>(ctx);  // residual rows -> L2 while the tile's GEMMs run
      for (int j = 0; j < Cfg::NCH; ++j, ++cnt) {
        const uint32_t buf = cnt & 1, use = cnt >> 1;
        SKY_T(0, mbar_wait(&acc1_full[buf], use & 1));
        SKY_T(1, mbar_wait(&hid_empty[buf], (use & 1) ^ 1));
        long long _tg = dbg ? clock64() : 0;
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * Cfg::HC + part * COLS_PER_WARP;
        uint8_t* hbuf = hid_s + buf * Cfg::HID_BYTES;
        const uint32_t r = q * 32 + lane;
#pragma unroll
        for (int g = 0; g < COLS_PER_WARP / 32; ++g) {
          float v[32];
          SKY_T(6, tmem_ld32(taddr + g * 32, v));
          const int hc = part * COLS_PER_WARP + g * 32;          // column inside the chunk
          const uint32_t bb = b1s_s + (j * Cfg::HC + hc) * 4;
#pragma unroll
          if (!(expflags & 1)) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b4 = lds_f32x4_ro(bb + i * 4);
            gelu_erf_x2(v[i], v[i + 1], b4.x, b4.y);
            gelu_erf_x2(v[i + 2], v[i + 3], b4.z, b4.w);
          }
          }
          uint8_t* kbase = hbuf + (hc >> 6) * G2_A_BYTES;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 pk;
            pk.x = pack_half2(v[8 * i], v[8 * i + 1]); pk.y = pack_half2(v[8 * i + 2], v[8 * i + 3]);
            pk.z = pack_half2(v[8 * i + 4], v[8 * i + 5]); pk.w = pack_half2(v[8 * i + 6], v[8 * i + 7]);
            *reinterpret_cast<uint4*>(kbase + sw128_offset(r, ((hc & 63) >> 3) + i)) = pk;
          }
        }
        if (dbg) tacc[2] += clock64() - _tg;
        long long _tf = dbg ? clock64() : 0;
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) { mbar_arrive(&acc1_empty[buf]); mbar_arrive(&hid_full[buf]); }
        if (dbg) tacc[3] += clock64() - _tf;
      }
      // ---- LayerNorm + residual on the finished acc2 tile ----
      SKY_T(4, mbar_wait(acc2_full, tph));
      tc_fence_after();
      long long _tl = dbg ? clock64() : 0;
      ctx.row0 = (long long)mt * 128 + q * 32;
      AccTmem2 acc{tmem_base + ((uint32_t)(q * 32) << 16) + Cfg::ACC2_COL};
      if (!(expflags & 2)) epi.template run<C>(acc, ctx);
      tc_fence_before();
      // the LN patches alias the hidden buffers: no epilogue warp may start the next tile's
      // GELU stores before every warp has left its patch
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (lane == 0) mbar_arrive(acc2_empty);
      if (dbg) tacc[5] += clock64() - _tl;
    }
  }
  if (dbg && lane == 0 && (warp == 0 || warp == 9) && blockIdx.x < 4) {
    for (int i = 0; i < 16; ++i) dbg[(blockIdx.x * 2 + (warp == 9)) * 16 + i] = tacc[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 9) {
    __syncwarp();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int C, class EpiT = EpiLnRes>
int launch_mlp_fused(const uint8_t* xh_in, const EpiT& epi, const uint8_t* W1img,
                     const uint8_t* W2img, const float* b1, long long M, int num_sms, cudaStream_t st) {
  static long long* dbg = nullptr;
  static int dbg_runs = 0;
  if (getenv("SKY_MLP_DBG") && !dbg) cudaMallocManaged(&dbg, 256 * 8);
  using Cfg = MlpCfg<C>;
  auto kern = k_mlp_fused<C, EpiT>;
  static std::atomic<uint64_t> configured{0};   // one bit per device: the attribute is per (function, device)
  if (int rc = smem_opt_in(configured, reinterpret_cast<const void*>(kern), Cfg::SMEM_BYTES)) return rc;
  const int tiles = (int)((M + 127) / 128);
  const int grid = tiles < num_sms ? tiles : num_sms;
  static const int expflags = getenv("SKY_MLP_EXP") ? atoi(getenv("SKY_MLP_EXP")) : 0;  // timing experiments only
  kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(xh_in, epi, W1img, W2img, b1, M, tiles, dbg, expflags);
  if (dbg && dbg_runs < 2) {
    ++dbg_runs;
    cudaDeviceSynchronize();
    for (int b = 0; b < 2; ++b) {
      const long long* e = dbg + (b * 2) * 16; const long long* m = dbg + (b * 2 + 1) * 16;
      printf("[mlp C=%d cta %d, %d tiles/cta] EPI wait_acc1 %lld wait_hidempty %lld gelu %lld (ldtm %lld) fence+arrive %lld wait_acc2 %lld ln %lld | "
             "MMA wait_a %lld wait_acc1empty %lld wait_w1 %lld wait_hidfull %lld wait_w2 %lld\n", C, b, (tiles + grid - 1) / grid,
             e[0], e[1], e[2], e[6], e[3], e[4], e[5], m[0], m[1], m[2], m[3], m[4]);
    }
  }
  SKY_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace sky
