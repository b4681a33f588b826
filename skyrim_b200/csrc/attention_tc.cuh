// Earth-specific 3-D window attention on tcgen05 + TMEM, fed by bulk (TMA-engine) copies.
//
//     O = softmax(scale * Q K^T + B[type, head] + shift_mask) V      per (window, head); 144 tokens, head_dim 32
//
// Data layout (written by EpiQkvWin, the QKV projection's epilogue, and k_qkv_fill_pad):
//   qkv "window image"  [part q|k|v][member][window][head pair] -> tile of 144 rows x 128 B
//     row j    = position in the (rolled) window, (zj, hj, wj) order
//     128 B    = [head 2p: 32 halves | head 2p+1: 32 halves], 16-byte chunk c stored at c ^ (j & 7)
//   i.e. byte for byte a K-major SWIZZLE_128B UMMA operand for Q and K (the two heads are k-steps {0,1} and {2,3}
//   of the 64-wide row) and an MN-major SWIZZLE_128B B operand for V ([key][dims]); tools/umma_probe.cu verified
//   every descriptor variant used here on the hardware (profiles/r2_umma_probe.md).  Windowing, the cyclic shift and
//   the latitude padding were applied by the producer, so one work item is FOUR contiguous bulk copies.
//
// One persistent CTA per SM, 320 threads:
//   warp 0     loader: cp.async.bulk global -> shared (Q rows 0..127, Q rows 128..143, K, V of one window / head pair)
//   warp 1     MMA issuer (one elected thread): S = Q K^T (M=128, N=144, K=32) into TMEM, O = P V with P read from
//              TMEM (A operand) and V from shared memory
//   warps 2..9 two softmax groups of four warps (TMEM lane quarter = warp % 4): thread = query row.  The 144 scores
//              of the row come out of TMEM into registers, the bias (expanded per (window type, head) into shared
//              memory, shift mask folded in) is added, exp2 / row sum, P goes back to TMEM as fp16 over the scores;
//              the same thread later scales its O row by 1/sum and stores 64 B of the projection's operand image.
// 144 query rows = one M=128 tile + 16 left-over rows.  The left-over rows of the TWO heads of a pair share one
// M=128 tile: head A's rows sit in lanes 32q..32q+15 and head B's in lanes 32q+16..32q+31 of a single accumulator
// (two accumulating MMA pairs whose A tiles start at different rows of [Q rows | 16 zero rows | left-over rows |
// 16 zero rows]), with q rotating over the lane quarters from item to item, so all softmax lanes stay busy:
// 9 warp-tasks of 32 valid rows per window / head pair.
#pragma once
#include "gemm2.cuh"

namespace sky {

constexpr int AT_TABLE = (2 * WW - 1) * WH * WH * WZ * WZ;   // 3312 entries of the compact bias table per (type, head)
constexpr int AT_TILE_B = WIN_TOK * 128;                     // 18432: one (window, head pair) tile of q, k or v
constexpr int AT_BIAS_LD = 304;                              // bytes per expanded-bias row: 144 halves + 16 B pad (conflict-free LDS.128)
constexpr int AT_BIAS_HEAD_B = WIN_TOK * AT_BIAS_LD;         // 43776
constexpr int AT_BIAS_B = (2 * AT_BIAS_HEAD_B + 1023) / 1024 * 1024;   // 88064
constexpr int AT_Q = 0, AT_Z0 = 128 * 128, AT_L = AT_Z0 + 2048, AT_Z1 = AT_L + 2048, AT_K = AT_Z1 + 2048, AT_V = AT_K + AT_TILE_B;
constexpr int AT_STAGE_B = AT_V + AT_TILE_B;                 // 59392
constexpr int AT_STAGES = 2;
constexpr int AT_ITEM_TX = 128 * 128 + 2048 + 2 * AT_TILE_B; // bytes landing per item
constexpr int AT_SMEM_BYTES = 1024 + AT_BIAS_B + AT_STAGES * AT_STAGE_B + 256;
constexpr int AT_THREADS = 320;
constexpr int AT_S_COLS = 144, AT_O_COL0 = 2 * AT_S_COLS, AT_O_COLS = 64;   // TMEM: S0 | S1 | O0 | O1 = 416 of 512 columns
static_assert(AT_STAGE_B % 1024 == 0 && AT_BIAS_B % 1024 == 0, "swizzle atom alignment");
static_assert(AT_SMEM_BYTES <= 232448, "smem budget");

struct AttnArgs {
  const uint8_t* qkv;        // window image, 3 parts
  long long part_stride;     // bytes between the q, k and v parts
  uint8_t* att_img; int att_nkb;   // output: fp16 tile image of (tokens, C) in natural token order
  const __half* bias_tab;    // (n_type, heads, 3312) fp16, pre-scaled by log2 e
  Geo g; int roll; int B; int pairs;
  float sl2;                 // head_dim^-0.5 * log2 e
  float mask_l2;             // mask value * log2 e
  long long items;           // B * nWin * pairs
};

// instruction descriptors: A, B fp16, D fp32; B MN-major for the PV products
__host__ __device__ constexpr uint32_t make_idesc_f16_bmn(int M, int N) { return make_idesc_f16(M, N) | (1u << 16); }

__device__ __forceinline__ void tc_mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint4 lds_b128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float2 h2_to_f2(uint32_t u) { return __half22float2(*reinterpret_cast<const __half2*>(&u)); }

// window position (wz, wh, ww, row j) -> natural token of member b, or -1 for a latitude-padding row
__device__ __forceinline__ long long at_row_token(const Geo& g, int roll, int b, int wz, int wh, int ww, int j) {
  const int wj = j % WW, hj = (j / WW) % WH, zj = j / (WW * WH);
  int z = wz * WZ + zj, h = wh * WH + hj, w = ww * WW + wj;
  if (roll) {
    z += SZ; if (z >= g.Z) z -= g.Z;
    h += SH; if (h >= g.Hp) h -= g.Hp;
    w += SW; if (w >= g.W) w -= g.W;
  }
  if (h >= g.H) return -1;
  return (long long)b * g.T + ((long long)z * g.H + h) * g.W + w;
}

// ---------------------------------------------------------------------------------------------------------------
// QKV projection epilogue: accumulator (tokens in natural order) -> fp16 window image.  Re-tiled through the warp's
// swizzled patch exactly like Epi2F16; only the destination address differs: row (window, j) of tile (part, pair).
// A warp store covers 8 token rows x 64 B (one head of each row).
// ---------------------------------------------------------------------------------------------------------------
struct EpiQkvWin {
  static constexpr bool kNeedsBias = false;
  uint8_t* out; long long part_stride; const float* bias; Geo g; int roll; int C; int pairs;
  const float* gamma = nullptr; const float* beta = nullptr;   // unused (uniform epilogue interface)
#ifdef SKY_EXPERIMENTS
  int exp = 0;
#else
  static constexpr int exp = 0;
#endif
  template <int BN, class Acc>
  __device__ void run(Acc& acc, const EpiCtx& x) const {
    // this lane's own token row -> index of its 128-byte row inside a part, pair 0 (tile stride = 144 rows)
    long long myrow = -1;
    {
      const long long row = x.row0 + x.lane;
      if (row < x.M) {
        const int t = (int)(row % g.T); const long long b = row / g.T;
        int w = t % g.W; const int q2 = t / g.W; int h = q2 % g.H, z = q2 / g.H;
        if (roll) {   // natural -> rolled-grid coordinates
          z -= SZ; if (z < 0) z += g.Z;
          h -= SH; if (h < 0) h += g.Hp;
          w -= SW; if (w < 0) w += g.W;
        }
        const int wz = z / WZ, zj = z % WZ, wh = h / WH, hj = h % WH, ww = w / WW, wj = w % WW;
        const long long win = b * g.nWin + ((long long)wz * g.nWh + wh) * g.nWw + ww;
        myrow = win * pairs * WIN_TOK + (zj * WH + hj) * WW + wj;
      }
    }
    const int rsub = x.lane >> 2, ch = x.lane & 3;
    for (int c = x.part * 32; c < BN; c += 32 * x.nparts) {
      const int col = x.n0 + c;                       // first of the 32 columns = one head of one part
      const int part = col / C, hd = (col - part * C) >> 5;
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + col + ch * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + col + ch * 8 + 4));
      {
        float v[32];
        acc.load32(c, v);
        patch_put_v(x.patch_s, x.lane, v);
      }
      __syncwarp();
      uint8_t* pbase = out + (size_t)part * part_stride + (size_t)(hd >> 1) * AT_TILE_B;
      const uint32_t cpos = (uint32_t)((hd & 1) * 4 + ch);
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int rr = it * 8 + rsub;
        float4 t0 = lds_f32x4(patchv_addr(x.patch_s, rr, 2 * ch));
        float4 t1 = lds_f32x4(patchv_addr(x.patch_s, rr, 2 * ch + 1));
        t0.x += b0.x; t0.y += b0.y; t0.z += b0.z; t0.w += b0.w;
        t1.x += b1.x; t1.y += b1.y; t1.z += b1.z; t1.w += b1.w;
        uint4 pk;
        pk.x = pack_half2(t0.x, t0.y); pk.y = pack_half2(t0.z, t0.w);
        pk.z = pack_half2(t1.x, t1.y); pk.w = pack_half2(t1.z, t1.w);
        const long long drow = __shfl_sync(0xffffffffu, myrow, rr);
        if (drow >= 0 && !(exp & 2))
          *reinterpret_cast<uint4*>(pbase + (size_t)drow * 128 + ((cpos ^ ((uint32_t)drow & 7u)) << 4)) = pk;
      }
      __syncwarp();
    }
  }
};

// latitude-padding rows of the window image (tokens with x = 0): q, k, v = the projection's bias
__global__ void __launch_bounds__(256) k_qkv_fill_pad(uint8_t* __restrict__ out, long long part_stride, const float* __restrict__ bias,
                                                      Geo g, int roll, int B, int pairs, int C, long long total) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int chunk = (int)(t & 7); t >>= 3;
  const int pair = (int)(t % pairs); t /= pairs;
  int w = (int)(t % g.W); t /= g.W;
  const int npad = g.Hp - g.H;
  int h = g.H + (int)(t % npad); t /= npad;
  int z = (int)(t % g.Z); t /= g.Z;
  const int b = (int)(t % B); const int part = (int)(t / B);
  if (roll) {
    z -= SZ; if (z < 0) z += g.Z;
    h -= SH; if (h < 0) h += g.Hp;
    w -= SW; if (w < 0) w += g.W;
  }
  const int wz = z / WZ, zj = z % WZ, wh = h / WH, hj = h % WH, ww = w / WW, wj = w % WW;
  const long long win = (long long)b * g.nWin + ((long long)wz * g.nWh + wh) * g.nWw + ww;
  const int j = (zj * WH + hj) * WW + wj;
  const float* bp = bias + part * C + pair * 64 + chunk * 8;
  uint4 pk;
  pk.x = pack_half2(bp[0], bp[1]); pk.y = pack_half2(bp[2], bp[3]);
  pk.z = pack_half2(bp[4], bp[5]); pk.w = pack_half2(bp[6], bp[7]);
  *reinterpret_cast<uint4*>(out + (size_t)part * part_stride + ((size_t)(win * pairs + pair) * WIN_TOK + j) * 128 +
                            (((uint32_t)chunk ^ ((uint32_t)j & 7u)) << 4)) = pk;
}

// ---------------------------------------------------------------------------------------------------------------
// the attention kernel
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(AT_THREADS, 1) k_window_attention_tc(const AttnArgs a) {
  extern __shared__ uint8_t at_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bias_s = smem;
  uint8_t* stages = smem + AT_BIAS_B;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stages + AT_STAGES * AT_STAGE_B);
  uint64_t* full = bars;            // [2] loader -> MMA (bulk-copy bytes)
  uint64_t* empty = bars + 2;       // [2] MMA -> loader (tcgen05.commit after the item's last PV product)
  uint64_t* s_full = bars + 4;      // [2] MMA -> softmax group: scores of a tile are in TMEM
  uint64_t* p_ready = bars + 6;     // [2] softmax group (4 warps) -> MMA: P is in TMEM, the group's O buffer is drained
  uint64_t* o_full = bars + 8;      // [2] MMA -> softmax group: O of a tile is complete (waited for only at a segment's end)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x / 32, 0), lane = threadIdx.x % 32;
  const int tid = threadIdx.x;
  const Geo& g = a.g;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 1); mbar_init(&empty[i], 1); mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], 4); mbar_init(&o_full[i], 1);
    }
    mbar_fence_init();
  }
  // the two 16-row zero blocks around the left-over rows of each stage: written once, never overwritten
  for (int i = tid; i < AT_STAGES * 2 * 128; i += AT_THREADS) {
    const int st = i / 256, r = i % 256;
    *reinterpret_cast<uint4*>(stages + st * AT_STAGE_B + (r < 128 ? AT_Z0 : AT_Z1 - 2048) + r * 16) = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async_smem();
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;

  // this CTA's contiguous share of the items; item n = ((type * pairs + pair) * B + b) * nWw + ww
  const long long n_begin = a.items * blockIdx.x / gridDim.x, n_end = a.items * (blockIdx.x + 1) / gridDim.x;
  const int per_group = a.B * g.nWw;

  uint32_t ld_items = 0;            // loader / MMA: items of all segments so far (stage = n & 1, phase = (n >> 1) & 1)
  uint32_t grp_uses[2] = {0, 0};    // MMA: p_ready waits per group; softmax warps use [0] for their own group

  for (long long seg0 = n_begin; seg0 < n_end;) {
    const long long key = seg0 / per_group;
    long long seg1 = (key + 1) * per_group;
    if (seg1 > n_end) seg1 = n_end;
    const int n_items = (int)(seg1 - seg0);
    const int type = (int)(key / a.pairs), pair = (int)(key % a.pairs);
    const int wz = type / g.nWh, wh = type % g.nWh;

    // ---- expand the bias of the pair's two heads: B[i][j] = table[idx(i, j)] (+ shift mask), fp16, log2 units ----
    {
      const int fmask = a.roll ? ((wz == g.nWz - 1 ? 1 : 0) | (wh == g.nWh - 1 ? 2 : 0)) : 0;
      for (int u = tid; u < 2 * WIN_TOK * 18; u += AT_THREADS) {
        const int hd = u / (WIN_TOK * 18), r = u % (WIN_TOK * 18), i = r / 18, c8 = r % 18;
        const int wi = i % WW, hi = (i / WW) % WH, zi = i / (WW * WH);
        const int rowpart = zi * ((2 * WW - 1) * WH * WH) + hi * (2 * WW - 1) + wi + (WW - 1);
        const int rflag = (zi >= WZ - SZ ? 1 : 0) | (hi >= WH - SH ? 2 : 0);
        const __half* tab = a.bias_tab + ((long long)type * g.heads + 2 * pair + hd) * AT_TABLE;
        __half v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int j = c8 * 8 + e;
          const int wj = j % WW, hj = (j / WW) % WH, zj = j / (WW * WH);
          const int idx = rowpart + (WZ * zj) * ((2 * WW - 1) * WH * WH) + (WH * hj) * (2 * WW - 1) - wj;
          float f = __half2float(__ldg(tab + idx));
          const int cflag = (zj >= WZ - SZ ? 1 : 0) | (hj >= WH - SH ? 2 : 0);
          if ((rflag ^ cflag) & fmask) f += a.mask_l2;
          v[e] = __float2half_rn(f);
        }
        *reinterpret_cast<uint4*>(bias_s + hd * AT_BIAS_HEAD_B + i * AT_BIAS_LD + c8 * 16) = *reinterpret_cast<const uint4*>(v);
      }
    }
    __syncthreads();

    const int T = 3 * n_items;      // tiles of the segment: per item [head A rows 0..127, head B rows 0..127, left-over rows of both]
    if (warp == 0) {
      // ===================== loader =====================
      for (int it = 0; it < n_items; ++it) {
        const uint32_t n = ld_items + it, s = n & 1;
        mbar_wait(&empty[s], ((n >> 1) & 1) ^ 1);
        if (lane == 0) {
          const long long item = seg0 + it;
          const int ww = (int)(item % g.nWw), b = (int)((item / g.nWw) % a.B);
          const uint8_t* src = a.qkv + ((size_t)(((long long)b * g.nWin + (long long)type * g.nWw + ww) * a.pairs + pair)) * AT_TILE_B;
          uint8_t* dst = stages + s * AT_STAGE_B;
          mbar_arrive_expect_tx(&full[s], AT_ITEM_TX);
          bulk_g2s(dst + AT_Q, src, 128 * 128, &full[s]);
          bulk_g2s(dst + AT_L, src + 128 * 128, 2048, &full[s]);
          bulk_g2s(dst + AT_K, src + a.part_stride, AT_TILE_B, &full[s]);
          bulk_g2s(dst + AT_V, src + 2 * a.part_stride, AT_TILE_B, &full[s]);
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc_s = make_idesc_f16(128, AT_S_COLS);
      constexpr uint32_t idesc_o32 = make_idesc_f16_bmn(128, 32), idesc_o64 = make_idesc_f16_bmn(128, 64);
      const uint32_t stage0 = smem_u32(stages);
      auto issue_scores = [&](int t) {
        const int it = t / 3, kind = t - 3 * it, grp = t & 1;
        const uint32_t n = ld_items + it, sb = stage0 + (n & 1) * AT_STAGE_B;
        if (kind == 0) { mbar_wait(&full[n & 1], (n >> 1) & 1); tc_fence_after(); }
        const uint32_t d = tmem + grp * AT_S_COLS;
        const uint64_t dk = make_desc_sw128(sb + AT_K);
        if (kind < 2) {
          const uint64_t dq = make_desc_sw128(sb + AT_Q) + 4 * kind;   // head B = k-steps 2, 3 of the 64-wide row (+64 B)
          if (elect_one()) {
            tc_mma_f16(d, dq, dk + 4 * kind, idesc_s, 0u);
            tc_mma_f16(d, dq + 2, dk + 4 * kind + 2, idesc_s, 1u);
            tc_commit(&s_full[grp]);
          }
        } else {
          // left-over rows of both heads into lanes 32q .. 32q+31:  [head A rows | zeros] x K_A  +  [zeros | head B rows] x K_B
          const uint32_t q = (uint32_t)((seg0 + it) & 3);
          const uint64_t dlo = make_desc_sw128(sb + AT_L - q * 4096), dhi = make_desc_sw128(sb + AT_Z0 - q * 4096);
          if (elect_one()) {
            tc_mma_f16(d, dlo, dk, idesc_s, 0u);
            tc_mma_f16(d, dlo + 2, dk + 2, idesc_s, 1u);
            tc_mma_f16(d, dhi + 4, dk + 4, idesc_s, 1u);
            tc_mma_f16(d, dhi + 6, dk + 6, idesc_s, 1u);
            tc_commit(&s_full[grp]);
          }
        }
        __syncwarp();
      };
      issue_scores(0);
      if (T > 1) issue_scores(1);
      for (int t = 0; t < T; ++t) {
        const int it = t / 3, kind = t - 3 * it, grp = t & 1;
        const uint32_t n = ld_items + it, sb = stage0 + (n & 1) * AT_STAGE_B;
        mbar_wait(&p_ready[grp], grp_uses[grp] & 1);
        ++grp_uses[grp];
        tc_fence_after();
        {
          const uint32_t o = tmem + AT_O_COL0 + grp * AT_O_COLS, pa = tmem + grp * AT_S_COLS;
          const uint64_t dv = make_desc_sw128(sb + AT_V) + (kind == 1 ? 4 : 0);   // head B's dims start 64 B into the row
          const uint32_t idesc = kind == 2 ? idesc_o64 : idesc_o32;
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 9; ++ks)    // 16 keys per step: 8 TMEM columns of P, 16 rows (2048 B) of V
              tc_mma_f16_ts(o, pa + 8 * ks, dv + 128 * ks, idesc, ks != 0 ? 1u : 0u);
            tc_commit(&o_full[grp]);
            if (kind == 2) tc_commit(&empty[n & 1]);
          }
          __syncwarp();
        }
        if (t + 2 < T) issue_scores(t + 2);
      }
    } else {
      // ===================== softmax groups =====================
      const int grp = (warp - 2) >> 2, q = warp & 3;
      const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
      const uint32_t s_t = lane_base + grp * AT_S_COLS, o_t = lane_base + AT_O_COL0 + grp * AT_O_COLS;
      bool have_prev = false, prev_active = false;
      int prev_kind = 0; long long prev_item = 0; float prev_inv = 0.f;
      auto epilogue = [&]() {
        if (!prev_active) return;
        // row of the window and head handled by this thread in the previous tile
        const int i = prev_kind < 2 ? q * 32 + lane : 128 + (lane & 15);
        const int hd = prev_kind < 2 ? prev_kind : (lane >> 4);
        float o[32];
        if (prev_kind < 2) {
          tmem_ld32(o_t, o);
        } else {
          float o2[32];
          tmem_ld32_nowait(o_t, o);
          tmem_ld32_nowait(o_t + 32, o2);
          tmem_ld_wait();
          if (lane >= 16) {
#pragma unroll
            for (int e = 0; e < 32; ++e) o[e] = o2[e];
          }
        }
        const int ww = (int)(prev_item % g.nWw), b = (int)((prev_item / g.nWw) % a.B);
        const long long tok = at_row_token(g, a.roll, b, wz, wh, ww, i);
        if (tok >= 0) {
          uint8_t* row = a.att_img + ((size_t)(tok >> 7) * a.att_nkb + pair) * (size_t)G2_A_BYTES + (size_t)(tok & 127) * 128;
          const uint32_t r7 = (uint32_t)tok & 7u;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint4 pk;
            pk.x = pack_half2(o[8 * c] * prev_inv, o[8 * c + 1] * prev_inv); pk.y = pack_half2(o[8 * c + 2] * prev_inv, o[8 * c + 3] * prev_inv);
            pk.z = pack_half2(o[8 * c + 4] * prev_inv, o[8 * c + 5] * prev_inv); pk.w = pack_half2(o[8 * c + 6] * prev_inv, o[8 * c + 7] * prev_inv);
            *reinterpret_cast<uint4*>(row + ((((uint32_t)(hd * 4 + c)) ^ r7) << 4)) = pk;
          }
        }
      };
      for (int t = grp; t < T; t += 2) {
        const int it = t / 3, kind = t - 3 * it;
        const long long item = seg0 + it;
        mbar_wait(&s_full[grp], grp_uses[0] & 1);
        ++grp_uses[0];
        tc_fence_after();
        if (have_prev) epilogue();   // the previous tile's O is complete: its PV product was issued before this tile's QK^T
        const bool active = kind < 2 || q == (int)(item & 3);
        float inv = 0.f;
        if (active) {
          const int i = kind < 2 ? q * 32 + lane : 128 + (lane & 15);
          const int hd = kind < 2 ? kind : (lane >> 4);
          const uint32_t brow = smem_u32(bias_s) + hd * AT_BIAS_HEAD_B + i * AT_BIAS_LD;
          float s[AT_S_COLS];
          __syncwarp();
          tmem_ld32_nowait(s_t, *reinterpret_cast<float(*)[32]>(s));
          tmem_ld32_nowait(s_t + 32, *reinterpret_cast<float(*)[32]>(s + 32));
          tmem_ld32_nowait(s_t + 64, *reinterpret_cast<float(*)[32]>(s + 64));
          tmem_ld32_nowait(s_t + 96, *reinterpret_cast<float(*)[32]>(s + 96));
          tmem_ld16_nowait(s_t + 128, s + 128);
          tmem_ld_wait();
          float mx = -INFINITY;
#pragma unroll
          for (int c8 = 0; c8 < 18; ++c8) {
            const uint4 bb = lds_b128(brow + c8 * 16);
            const float2 b0 = h2_to_f2(bb.x), b1 = h2_to_f2(bb.y), b2 = h2_to_f2(bb.z), b3 = h2_to_f2(bb.w);
            float* sp = s + 8 * c8;
            sp[0] = fmaf(sp[0], a.sl2, b0.x); sp[1] = fmaf(sp[1], a.sl2, b0.y);
            sp[2] = fmaf(sp[2], a.sl2, b1.x); sp[3] = fmaf(sp[3], a.sl2, b1.y);
            sp[4] = fmaf(sp[4], a.sl2, b2.x); sp[5] = fmaf(sp[5], a.sl2, b2.y);
            sp[6] = fmaf(sp[6], a.sl2, b3.x); sp[7] = fmaf(sp[7], a.sl2, b3.y);
            mx = fmaxf(mx, fmaxf(fmaxf(fmaxf(sp[0], sp[1]), fmaxf(sp[2], sp[3])), fmaxf(fmaxf(sp[4], sp[5]), fmaxf(sp[6], sp[7]))));
          }
          // P = exp2(s - max) in (0, 1]: un-normalised into the tensor core, 1/sum applied to the 32 outputs of the row.
          // Written back over the scores 32 keys (16 packed columns) at a time: every score is already in registers.
          float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
          for (int c = 0; c < 5; ++c) {
            uint32_t p[16];
#pragma unroll
            for (int k = 0; k < (c < 4 ? 16 : 8); ++k) {
              const float e0 = mufu_ex2(s[32 * c + 2 * k] - mx), e1 = mufu_ex2(s[32 * c + 2 * k + 1] - mx);
              sum0 += e0; sum1 += e1;
              p[k] = pack_half2(e0, e1);
            }
            if (c < 4) tmem_st16(s_t + 16 * c, p);
            else tmem_st8(s_t + 64, p);
          }
          inv = 1.f / (sum0 + sum1);
          tmem_st_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[grp]);
        have_prev = true; prev_active = active; prev_kind = kind; prev_item = item; prev_inv = inv;
      }
      if (have_prev) {
        mbar_wait(&o_full[grp], (grp_uses[0] - 1) & 1);
        tc_fence_after();
        epilogue();
      }
    }
    ld_items += (uint32_t)n_items;
    tc_fence_before();
    __syncthreads();   // all of this segment's reads of the bias are done; TMEM buffers are free
    tc_fence_after();
    seg0 = seg1;
  }
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc<512>(tmem);
  }
}

#ifdef SKY_EXPERIMENTS
// CUDA-core reference of the same operator on the same window image (development library, SKY_ATTN=ref): one thread
// per query row, fp32 scores, bias gathered from the compact table (independent of the expansion above).  Bisects
// "window image / index arithmetic" from "tensor pipeline" failures; never benchmarked.
__global__ void __launch_bounds__(WIN_TOK) k_window_attention_ref(const AttnArgs a) {
  const Geo& g = a.g;
  const long long item = blockIdx.x;
  const int hd = blockIdx.y;   // 0 / 1 of the pair
  const long long key = item / (a.B * g.nWw);
  const int type = (int)(key / a.pairs), pair = (int)(key % a.pairs), wz = type / g.nWh, wh = type % g.nWh;
  const int ww = (int)(item % g.nWw), b = (int)((item / g.nWw) % a.B);
  const uint8_t* src = a.qkv + ((size_t)(((long long)b * g.nWin + (long long)type * g.nWw + ww) * a.pairs + pair)) * AT_TILE_B;
  __shared__ float ks[WIN_TOK][33], vs[WIN_TOK][33];
  const int i = threadIdx.x;
  auto elem = [&](int part, int row, int d) {
    const uint32_t c = (uint32_t)(hd * 4 + d / 8);
    return __half2float(*reinterpret_cast<const __half*>(src + (size_t)part * a.part_stride + (size_t)row * 128 + ((c ^ ((uint32_t)row & 7u)) << 4) + (d % 8) * 2));
  };
  float q[32];
  for (int d = 0; d < 32; ++d) { q[d] = elem(0, i, d); ks[i][d] = elem(1, i, d); vs[i][d] = elem(2, i, d); }
  __syncthreads();
  const int fmask = a.roll ? ((wz == g.nWz - 1 ? 1 : 0) | (wh == g.nWh - 1 ? 2 : 0)) : 0;
  const int wi = i % WW, hi = (i / WW) % WH, zi = i / (WW * WH);
  const int rflag = (zi >= WZ - SZ ? 1 : 0) | (hi >= WH - SH ? 2 : 0);
  const __half* tab = a.bias_tab + ((long long)type * g.heads + 2 * pair + hd) * AT_TABLE;
  float sc[WIN_TOK], mx = -INFINITY;
  for (int j = 0; j < WIN_TOK; ++j) {
    float d = 0.f;
    for (int e = 0; e < 32; ++e) d += q[e] * ks[j][e];
    const int wj = j % WW, hj = (j / WW) % WH, zj = j / (WW * WH);
    const int idx = (zi + WZ * zj) * ((2 * WW - 1) * WH * WH) + (hi + WH * hj) * (2 * WW - 1) + wi - wj + WW - 1;
    const int cflag = (zj >= WZ - SZ ? 1 : 0) | (hj >= WH - SH ? 2 : 0);
    d = d * a.sl2 + __half2float(tab[idx]) + (((rflag ^ cflag) & fmask) ? a.mask_l2 : 0.f);
    sc[j] = d; mx = fmaxf(mx, d);
  }
  float sum = 0.f, o[32];
  for (int e = 0; e < 32; ++e) o[e] = 0.f;
  for (int j = 0; j < WIN_TOK; ++j) {
    const float p = exp2f(sc[j] - mx);
    sum += p;
    const float ph = __half2float(__float2half_rn(p));
    for (int e = 0; e < 32; ++e) o[e] += ph * vs[j][e];
  }
  const long long tok = at_row_token(g, a.roll, b, wz, wh, ww, i);
  if (tok < 0) return;
  uint8_t* row = a.att_img + ((size_t)(tok >> 7) * a.att_nkb + pair) * (size_t)G2_A_BYTES + (size_t)(tok & 127) * 128;
  for (int e = 0; e < 32; ++e)
    *reinterpret_cast<__half*>(row + ((((uint32_t)(hd * 4 + e / 8)) ^ ((uint32_t)tok & 7u)) << 4) + (e % 8) * 2) = __float2half_rn(o[e] / sum);
}
#endif

inline int launch_window_attention_tc(const AttnArgs& a, int num_sms, cudaStream_t st) {
  static std::atomic<uint64_t> configured{0};
  if (int rc = smem_opt_in(configured, reinterpret_cast<const void*>(k_window_attention_tc), AT_SMEM_BYTES)) return rc;
  const long long grid = a.items < num_sms ? a.items : num_sms;
  k_window_attention_tc<<<(unsigned)grid, AT_THREADS, AT_SMEM_BYTES, st>>>(a);
  SKY_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace sky
