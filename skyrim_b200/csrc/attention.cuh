// Earth-specific 3-D window attention core for one (window, head):
//     O = softmax(scale * Q K^T + B[type, head] + shift_mask) V
// 144 tokens per window, head_dim 32.  Q/K/V arrive as fp16 (QKV GEMM epilogue), the
// bias is gathered from the compact (type, head, 3312) table by index arithmetic — the
// 144x144 expanded bias the ONNX graph carries (~1 GB) is never materialised.
// Scores, softmax and the accumulators are fp32; the two small GEMMs run on mma.sync
// m16n8k16 (they are ~6-11 % of the step's FLOPs; the big contractions are tcgen05).
#pragma once
#include "gemm2.cuh"

namespace sky {

constexpr int ATT_THREADS = 96;
constexpr int ATT_LDS = 40;  // halves per smem row (32 + 8 pad): conflict-free ldmatrix
constexpr int ATT_TABLE = (2 * WW - 1) * WH * WH * WZ * WZ;  // 3312
constexpr int ATT_SMEM_BYTES = 3 * WIN_TOK * ATT_LDS * 2 + ATT_TABLE * 2 + 2 * WIN_TOK * 4 + WIN_TOK * 8;  // 43.5 KB: 5 CTAs per SM

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// grid (heads, windows, members).  qkv: (3, heads, tokens, 32) fp16, tokens in NATURAL order — the window
// partition, the cyclic shift and the latitude padding are index arithmetic here and nowhere
// else; a padding token has x = 0, so its q/k/v are the QKV bias.  Output: fp16 tile image of
// (tokens, C) (A operand of the projection GEMM), again in natural order.
// 5 CTAs / SM: the register file allocates 512 registers per warp-granule, so 130 registers cost a whole CTA of occupancy
__global__ void __launch_bounds__(ATT_THREADS, 5)
k_window_attention(const __half* __restrict__ qkv, uint8_t* __restrict__ att_img, int att_nkb,
                   const __half* __restrict__ bias_tab, const float* __restrict__ qkv_bias, Geo g, int roll,
                   float scale, float mask_value, long long rows /* tokens of all stacked members */) {
  extern __shared__ __align__(16) uint8_t att_smem[];
  __half* Qs = reinterpret_cast<__half*>(att_smem);
  __half* Ks = Qs + WIN_TOK * ATT_LDS;
  __half* Vs = Ks + WIN_TOK * ATT_LDS;
  __half* Bs = Vs + WIN_TOK * ATT_LDS;  // fp16 table slice (values ~0.03: rounding 1.5e-5 abs)
  int* colpart = reinterpret_cast<int*>(Bs + ATT_TABLE);  // per key token: column part of the table index (+64)
  int* colflag = colpart + WIN_TOK;                       // per key token: seam side flags (z: 1, lat: 2)
  long long* tok = reinterpret_cast<long long*>(colflag + WIN_TOK);

  // grid (heads, windows of one member, members): no 64-bit division by run-time values anywhere in the kernel
  const int head = blockIdx.x;
  const int win = blockIdx.y;
  const unsigned type = (unsigned)win / (unsigned)g.nWw;   // (z-window, lat-window)
  const int wzi = (int)(type / (unsigned)g.nWh), whi = (int)type - wzi * g.nWh;
  const int C = g.C;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;

  // natural token of window row j (same mapping as win_row_to_token, with the window coordinates — 64-bit divisions by
  // run-time values — taken out of the per-token loop: they were ~15 % of the kernel's instructions)
  const int wwi = win - (int)type * g.nWw;
  const long long mbase = (long long)blockIdx.z * g.T;
  for (int j = tid; j < WIN_TOK; j += ATT_THREADS) {
    int wj = j % WW, hj = (j / WW) % WH, zj = j / (WW * WH);
    {
      int z = wzi * WZ + zj, h = whi * WH + hj, w = wwi * WW + wj;
      if (roll) {
        z += SZ; if (z >= g.Z) z -= g.Z;
        h += SH; if (h >= g.Hp) h -= g.Hp;
        w += SW; if (w >= g.W) w -= g.W;
      }
      tok[j] = h >= g.H ? -1 : mbase + ((long long)z * g.H + h) * g.W + w;
    }
    int part = (WZ * zj) * ((2 * WW - 1) * WH * WH) + (WH * hj) * (2 * WW - 1) - wj;
    colpart[j] = part + 64;  // +64 keeps it non-negative (the row part carries -64)
    colflag[j] = (zj >= WZ - SZ ? 1 : 0) | (hj >= WH - SH ? 2 : 0);
  }
  const __half* bsrc = bias_tab + ((long long)type * g.heads + head) * ATT_TABLE;
  for (int i = tid; i < ATT_TABLE / 8; i += ATT_THREADS) cp_async16(Bs + 8 * i, bsrc + 8 * i);
  __syncthreads();
  // ---- stage Q, K, V (64 B per token each) ----
  // 4 threads per token row, 24 rows per pass: with the (part, head, token, 32) layout the 12 tokens of a window row are
  // 768 contiguous bytes, and every address below is one 64-bit multiply-add on per-thread constants
  {
    const int ch = tid & 3, rsub = tid >> 2;
    const size_t pstride = (size_t)g.heads * rows * 32;
    const __half* src0 = qkv + (size_t)head * rows * 32 + ch * 8;
#pragma unroll
    for (int r0 = 0; r0 < WIN_TOK; r0 += ATT_THREADS / 4) {
      const int row = r0 + rsub;
      const long long t = tok[row];
#pragma unroll
      for (int part = 0; part < 3; ++part) {
        __half* dst = Qs + (part * WIN_TOK + row) * ATT_LDS + ch * 8;
        if (t >= 0) {
          cp_async16(dst, src0 + part * pstride + t * 32);
        } else {
          const float* bq = qkv_bias + part * C + head * 32 + ch * 8;
          uint4 pk;
          pk.x = pack_half2(bq[0], bq[1]); pk.y = pack_half2(bq[2], bq[3]);
          pk.z = pack_half2(bq[4], bq[5]); pk.w = pack_half2(bq[6], bq[7]);
          *reinterpret_cast<uint4*>(dst) = pk;
        }
      }
    }
  }
  asm volatile("cp.async.commit_group;");
  asm volatile("cp.async.wait_group 0;");
  __syncthreads();

  const bool mz = roll && (wzi == g.nWz - 1);
  const bool mh = roll && (whi == g.nWh - 1);
  const int fmask = (mz ? 1 : 0) | (mh ? 2 : 0);  // CTA-uniform: only seam windows of shifted blocks mask
  const float sl2 = scale * 1.4426950408889634f;   // scores in log2 units: softmax via exp2
  const float mask_l2 = mask_value * 1.4426950408889634f;  // (the bias table is pre-scaled by log2 e)

  for (int rb = warp; rb < WIN_TOK / 16; rb += ATT_THREADS / 32) {
    const int r0 = rb * 16;
    // Q fragments for the two k-steps (d 0..15, 16..31)
    uint32_t qa[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
      ldsm_x4(qa[ks], Qs + (r0 + (lane & 15)) * ATT_LDS + ks * 16 + (lane >> 4) * 8);

    float s[18][4];
#pragma unroll
    for (int nt = 0; nt < 18; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      uint32_t kb[4];
      ldsm_x4(kb, Ks + (nt * 8 + (lane & 7)) * ATT_LDS + (lane >> 3) * 8);
      mma16816(s[nt], qa[0], kb[0], kb[1]);
      mma16816(s[nt], qa[1], kb[2], kb[3]);
    }
    // rows owned by this lane: i0 = r0 + lane/4, i1 = i0 + 8
    const int i0 = r0 + (lane >> 2), i1 = i0 + 8;
    int rowpart[2], rflag[2];
    {
      int wi = i0 % WW, hi = (i0 / WW) % WH, zi = i0 / (WW * WH);
      rowpart[0] = zi * ((2 * WW - 1) * WH * WH) + hi * (2 * WW - 1) + wi + (WW - 1) - 64;
      rflag[0] = (zi >= WZ - SZ ? 1 : 0) | (hi >= WH - SH ? 2 : 0);
      wi = i1 % WW; hi = (i1 / WW) % WH; zi = i1 / (WW * WH);
      rowpart[1] = zi * ((2 * WW - 1) * WH * WH) + hi * (2 * WW - 1) + wi + (WW - 1) - 64;
      rflag[1] = (zi >= WZ - SZ ? 1 : 0) | (hi >= WH - SH ? 2 : 0);
    }
    const __half* B0 = Bs + rowpart[0];
    const __half* B1 = Bs + rowpart[1];
    float mx0 = -INFINITY, mx1 = -INFINITY;
    if (fmask == 0) {
      // packed fp32x2 (FFMA2): the kernel is issue bound on this loop and the exp loop below
      const uint64_t sl22 = pack_f32x2(sl2, sl2);
#pragma unroll
      for (int nt = 0; nt < 18; ++nt) {
        const int2 cp2 = *reinterpret_cast<const int2*>(colpart + nt * 8 + 2 * (lane & 3));
        unpack_f32x2(fma_f32x2(pack_f32x2(s[nt][0], s[nt][1]), sl22, pack_f32x2(__half2float(B0[cp2.x]), __half2float(B0[cp2.y]))), s[nt][0], s[nt][1]);
        unpack_f32x2(fma_f32x2(pack_f32x2(s[nt][2], s[nt][3]), sl22, pack_f32x2(__half2float(B1[cp2.x]), __half2float(B1[cp2.y]))), s[nt][2], s[nt][3]);
        mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
      }
    } else {
#pragma unroll
      for (int nt = 0; nt < 18; ++nt) {
        const int j = nt * 8 + 2 * (lane & 3);
        const int2 cp2 = *reinterpret_cast<const int2*>(colpart + j);
        const int f0 = colflag[j], f1 = colflag[j + 1];
        s[nt][0] = fmaf(s[nt][0], sl2, __half2float(B0[cp2.x])) + (((rflag[0] ^ f0) & fmask) ? mask_l2 : 0.f);
        s[nt][1] = fmaf(s[nt][1], sl2, __half2float(B0[cp2.y])) + (((rflag[0] ^ f1) & fmask) ? mask_l2 : 0.f);
        s[nt][2] = fmaf(s[nt][2], sl2, __half2float(B1[cp2.x])) + (((rflag[1] ^ f0) & fmask) ? mask_l2 : 0.f);
        s[nt][3] = fmaf(s[nt][3], sl2, __half2float(B1[cp2.y])) + (((rflag[1] ^ f1) & fmask) ? mask_l2 : 0.f);
        mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float sum0, sum1;
    {
      const uint64_t nm0 = pack_f32x2(-mx0, -mx0), nm1 = pack_f32x2(-mx1, -mx1);
      uint64_t acc0 = pack_f32x2(0.f, 0.f), acc1 = acc0;
#pragma unroll
      for (int nt = 0; nt < 18; ++nt) {
        float d0, d1, d2, d3;
        unpack_f32x2(add_f32x2(pack_f32x2(s[nt][0], s[nt][1]), nm0), d0, d1);
        unpack_f32x2(add_f32x2(pack_f32x2(s[nt][2], s[nt][3]), nm1), d2, d3);
        s[nt][0] = mufu_ex2(d0); s[nt][1] = mufu_ex2(d1);
        s[nt][2] = mufu_ex2(d2); s[nt][3] = mufu_ex2(d3);
        acc0 = add_f32x2(acc0, pack_f32x2(s[nt][0], s[nt][1]));
        acc1 = add_f32x2(acc1, pack_f32x2(s[nt][2], s[nt][3]));
      }
      float a, b;
      unpack_f32x2(acc0, a, b); sum0 = a + b;
      unpack_f32x2(acc1, a, b); sum1 = a + b;
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    const float inv0 = 1.f / sum0, inv1 = 1.f / sum1;

    // O = P V : k runs over the 144 keys in 9 steps of 16.  P = exp2(s - max) in (0, 1] goes to
    // the tensor cores un-normalised; the 1/sum is applied to the 16x32 output instead of the
    // 16x144 probabilities.
    float o[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 9; ++kk) {
      uint32_t pa[4];
      pa[0] = pack_half2(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack_half2(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack_half2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack_half2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t vb[4];
        ldsm_x4_t(vb, Vs + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * ATT_LDS + np * 16 + (lane >> 4) * 8);
        mma16816(o[2 * np], pa, vb[0], vb[1]);
        mma16816(o[2 * np + 1], pa, vb[2], vb[3]);
      }
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) { o[nt][0] *= inv0; o[nt][1] *= inv0; o[nt][2] *= inv1; o[nt][3] *= inv1; }
    const long long t0 = tok[i0], t1 = tok[i1];
    const int colb = head * 32 + 2 * (lane & 3);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      if (t0 >= 0) *reinterpret_cast<uint32_t*>(att_img + img_offset(t0, colb + nt * 8, att_nkb)) = pack_half2(o[nt][0], o[nt][1]);
      if (t1 >= 0) *reinterpret_cast<uint32_t*>(att_img + img_offset(t1, colb + nt * 8, att_nkb)) = pack_half2(o[nt][2], o[nt][3]);
    }
  }
}

}  // namespace sky
