// Earth-specific 3-D window attention core for one (window, head):
//     O = softmax(scale * Q K^T + B[type, head] + shift_mask) V
// 144 tokens per window, head_dim 32.  Q/K/V arrive as fp16 (QKV GEMM epilogue), the
// bias is gathered from the compact (type, head, 3312) table by index arithmetic — the
// 144x144 expanded bias the ONNX graph carries (~1 GB) is never materialised.
// Scores, softmax and the accumulators are fp32; the two small GEMMs run on mma.sync
// m16n8k16 (they are ~6-11 % of the step's FLOPs; the big contractions are tcgen05).
#pragma once
#include "pangu_ops.cuh"

namespace sky {

constexpr int ATT_THREADS = 96;
constexpr int ATT_LDS = 40;  // halves per smem row (32 + 8 pad): conflict-free ldmatrix
constexpr int ATT_TABLE = (2 * WW - 1) * WH * WH * WZ * WZ;  // 3312
constexpr int ATT_SMEM_BYTES = 3 * WIN_TOK * ATT_LDS * 2 + ATT_TABLE * 4 + WIN_TOK * 4;

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

// grid (heads, total windows); qkv (rows, 3C) fp16 window-ordered; out (rows, C) fp16
__global__ void __launch_bounds__(ATT_THREADS)
k_window_attention(const __half* __restrict__ qkv, __half* __restrict__ out,
                   const float* __restrict__ bias_tab, Geo g, int roll, float scale, float mask_value) {
  extern __shared__ __align__(16) uint8_t att_smem[];
  __half* Qs = reinterpret_cast<__half*>(att_smem);
  __half* Ks = Qs + WIN_TOK * ATT_LDS;
  __half* Vs = Ks + WIN_TOK * ATT_LDS;
  float* Bs = reinterpret_cast<float*>(Vs + WIN_TOK * ATT_LDS);
  int* colinfo = reinterpret_cast<int*>(Bs + ATT_TABLE);  // per key token: table col-part | flags<<24

  const int head = blockIdx.x;
  const long long wing = blockIdx.y;           // global window index (members stacked)
  const int win = (int)(wing % g.nWin);
  const int type = win / g.nWw;                // (z-window, lat-window)
  const int whi = type % g.nWh, wzi = type / g.nWh;
  const int C = g.C;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;

  // ---- stage Q, K, V (64 B per row each) and the bias table slice ----
  const __half* src = qkv + wing * WIN_TOK * (3LL * C) + head * 32;
  for (int i = tid; i < WIN_TOK * 12; i += ATT_THREADS) {
    int row = i / 12, rem = i % 12, part = rem / 4, ch = rem % 4;
    cp_async16(Qs + part * WIN_TOK * ATT_LDS + row * ATT_LDS + ch * 8, src + (long long)row * 3 * C + part * C + ch * 8);
  }
  const float* bsrc = bias_tab + ((long long)type * g.heads + head) * ATT_TABLE;
  for (int i = tid; i < ATT_TABLE / 4; i += ATT_THREADS) cp_async16(Bs + 4 * i, bsrc + 4 * i);
  for (int j = tid; j < WIN_TOK; j += ATT_THREADS) {
    int wj = j % WW, hj = (j / WW) % WH, zj = j / (WW * WH);
    int part = (WZ * zj) * ((2 * WW - 1) * WH * WH) + (WH * hj) * (2 * WW - 1) - wj;
    int flags = (zj >= WZ - SZ ? 1 : 0) | (hj >= WH - SH ? 2 : 0);
    colinfo[j] = (part + 64) | (flags << 24);  // +64 keeps the packed field non-negative
  }
  asm volatile("cp.async.commit_group;");
  asm volatile("cp.async.wait_group 0;");
  __syncthreads();

  const bool mz = roll && (wzi == g.nWz - 1);
  const bool mh = roll && (whi == g.nWh - 1);
  const float sl2 = scale * 1.4426950408889634f;  // fold log2(e): softmax via exp2
  const float l2e = 1.4426950408889634f;

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
    const int fmask = (mz ? 1 : 0) | (mh ? 2 : 0);
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 18; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        int j = nt * 8 + 2 * (lane & 3) + e;
        int ci = colinfo[j];
        int cpart = ci & 0xffffff, cflag = ci >> 24;
        float b0 = Bs[rowpart[0] + cpart], b1 = Bs[rowpart[1] + cpart];
        float m0 = ((rflag[0] ^ cflag) & fmask) ? mask_value : 0.f;
        float m1 = ((rflag[1] ^ cflag) & fmask) ? mask_value : 0.f;
        // work in log2 units
        float v0 = s[nt][e] * sl2 + (b0 + m0) * l2e;
        float v1 = s[nt][2 + e] * sl2 + (b1 + m1) * l2e;
        s[nt][e] = v0; s[nt][2 + e] = v1;
        mx0 = fmaxf(mx0, v0); mx1 = fmaxf(mx1, v1);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 18; ++nt) {
      s[nt][0] = exp2f(s[nt][0] - mx0); s[nt][1] = exp2f(s[nt][1] - mx0);
      s[nt][2] = exp2f(s[nt][2] - mx1); s[nt][3] = exp2f(s[nt][3] - mx1);
      sum0 += s[nt][0] + s[nt][1];
      sum1 += s[nt][2] + s[nt][3];
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    const float inv0 = 1.f / sum0, inv1 = 1.f / sum1;

    // O = P V : k runs over the 144 keys in 9 steps of 16
    float o[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 9; ++kk) {
      uint32_t pa[4];
      pa[0] = pack_half2(s[2 * kk][0] * inv0, s[2 * kk][1] * inv0);
      pa[1] = pack_half2(s[2 * kk][2] * inv1, s[2 * kk][3] * inv1);
      pa[2] = pack_half2(s[2 * kk + 1][0] * inv0, s[2 * kk + 1][1] * inv0);
      pa[3] = pack_half2(s[2 * kk + 1][2] * inv1, s[2 * kk + 1][3] * inv1);
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t vb[4];
        ldsm_x4_t(vb, Vs + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * ATT_LDS + np * 16 + (lane >> 4) * 8);
        mma16816(o[2 * np], pa, vb[0], vb[1]);
        mma16816(o[2 * np + 1], pa, vb[2], vb[3]);
      }
    }
    __half* orow0 = out + (wing * WIN_TOK + i0) * (long long)C + head * 32 + 2 * (lane & 3);
    __half* orow1 = orow0 + 8LL * C;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      *reinterpret_cast<uint32_t*>(orow0 + nt * 8) = pack_half2(o[nt][0], o[nt][1]);
      *reinterpret_cast<uint32_t*>(orow1 + nt * 8) = pack_half2(o[nt][2], o[nt][3]);
    }
  }
}

}  // namespace sky
