// UMMA layout probe (development tool, not product code): runs single tcgen05.mma sequences with
// host-chosen descriptors / operand byte images and compares the TMEM accumulator with a CPU product.
// Settles, in one GPU call, the operand layouts the tcgen05 attention kernel relies on:
//   * K-major SWIZZLE_128B tiles with start offsets of whole 8-row groups and +32/+64 B K advances
//   * MN-major SWIZZLE_128B B operand (V as [key][dims]) with N = 32 at +0 / +64 B and N = 64
//   * A operand from TMEM (fp16 pairs per 32-bit column, row = lane)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe tools/umma_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <functional>

#include "../skyrim_b200/csrc/sky_common.cuh"

namespace sky { void set_error(const char*, ...) {} }
using namespace sky;

struct ProbeArgs {
  uint32_t a_bytes, b_bytes;          // operand images copied to smem (A at +0, B at +A_REGION)
  uint64_t desc_a_hi, desc_b_hi;      // descriptor bits [16, 64) (everything but the start address)
  uint32_t a_start, b_start;          // byte offset of the first k-step's start address inside its region
  uint32_t a_inc, b_inc;              // byte increment per k-step
  uint32_t idesc;
  int ksteps;
  int a_tmem;                         // 1: A comes from TMEM: a_words = [128][a_cols] uint32, column advance a_inc (columns) per k-step
  int a_cols;
  int d_cols;                         // accumulator columns to dump
};

constexpr int A_REGION = 64 * 1024;   // B region starts here
constexpr int SMEM_TOTAL = 1024 + 2 * 64 * 1024;

__device__ __forceinline__ void tc_mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}

__global__ void __launch_bounds__(128, 1)
k_probe(const uint8_t* __restrict__ a_img, const uint8_t* __restrict__ b_img, const uint32_t* __restrict__ a_words,
        float* __restrict__ d_out, ProbeArgs p) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid / 32;
  for (uint32_t i = tid * 16; i < p.a_bytes; i += 128 * 16) *reinterpret_cast<uint4*>(smem + i) = *reinterpret_cast<const uint4*>(a_img + i);
  for (uint32_t i = tid * 16; i < p.b_bytes; i += 128 * 16) *reinterpret_cast<uint4*>(smem + A_REGION + i) = *reinterpret_cast<const uint4*>(b_img + i);
  fence_proxy_async_smem();
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  if (p.a_tmem) {
    // each thread writes its own lane (row): a_cols words at columns 256..
    for (int c = 0; c < p.a_cols; c += 8) {
      const uint32_t* w = a_words + (size_t)tid * p.a_cols + c;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(lane_base + 256 + c),
                   "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                   : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + A_REGION);
      for (int k = 0; k < p.ksteps; ++k) {
        const uint64_t db = (p.desc_b_hi << 16) | (uint64_t)(((sb + p.b_start + k * p.b_inc) & 0x3FFFFu) >> 4);
        if (p.a_tmem) {
          tc_mma_f16_ts(tmem, tmem + 256 + k * p.a_inc, db, p.idesc, k != 0);
        } else {
          const uint64_t da = (p.desc_a_hi << 16) | (uint64_t)(((sa + p.a_start + k * p.a_inc) & 0x3FFFFu) >> 4);
          tc_mma_f16(tmem, da, db, p.idesc, k != 0);
        }
      }
      tc_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c = 0; c < p.d_cols; c += 32) {
    float v[32];
    tmem_ld32(lane_base + c, v);
    for (int j = 0; j < 32 && c + j < p.d_cols; ++j) d_out[(size_t)tid * p.d_cols + c + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------------
static uint64_t desc_hi(uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d >> 16;
}
static uint32_t idesc_f16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
static float frand() { return (float)((rand() % 2001) - 1000) / 1000.f; }
static uint16_t h16(float f) { __half h = __float2half_rn(f); uint16_t u; memcpy(&u, &h, 2); return u; }
static float f16(float f) { return __half2float(__float2half_rn(f)); }

struct Case {
  const char* name;
  int M, N, K;                       // logical product D[M, N] = A[M, K] B[N, K]^T over `ksteps * 16` of K
  std::vector<uint8_t> a_img, b_img;
  std::vector<uint32_t> a_words;
  ProbeArgs args;
  std::vector<float> A, B;           // logical matrices (fp16-rounded values), A[m*K+k], B[n*K+k]
  int row_lo = 0, row_hi = 128;      // rows of D that are meaningful
};

static void put16(std::vector<uint8_t>& img, size_t off, float v) {
  if (off + 2 > img.size()) img.resize(off + 2, 0);
  uint16_t u = h16(v);
  memcpy(&img[off], &u, 2);
}
// K-major SW128 tile: rows of 128 B (64 halves), k-blocks `rows*128` bytes apart
static size_t kmaj(int row, int k, int rows) { return (size_t)(k / 64) * rows * 128 + (size_t)row * 128 + (((k % 64) / 8) ^ (row & 7)) * 16 + (k % 8) * 2; }
// MN-major SW128 tile [k rows][128 B of n]
static size_t mnmaj(int k, int n) { return (size_t)k * 128 + (((n / 8) ^ (k & 7)) * 16) + (n % 8) * 2; }
// K-major SW64 tile: rows of 64 B (32 halves)
static size_t kmaj64(int row, int k) { return (size_t)row * 64 + ((((k % 32) / 8) ^ ((row >> 1) & 3)) * 16) + (k % 8) * 2; }

static int g_want = -1, g_idx = 0;
static int run(Case& c) {
  if (g_want >= 0 && g_idx++ != g_want) return 0;
  uint8_t *da, *db; uint32_t* dw; float* dd;
  c.a_img.resize((c.a_img.size() + 15) / 16 * 16 + 16, 0);
  c.b_img.resize((c.b_img.size() + 15) / 16 * 16 + 16, 0);
  if (c.a_words.empty()) c.a_words.resize(8, 0);
  cudaMalloc(&da, c.a_img.size()); cudaMalloc(&db, c.b_img.size()); cudaMalloc(&dw, c.a_words.size() * 4);
  cudaMalloc(&dd, 128 * 256 * 4);
  cudaMemcpy(da, c.a_img.data(), c.a_img.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(db, c.b_img.data(), c.b_img.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dw, c.a_words.data(), c.a_words.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dd, 0, 128 * 256 * 4);
  c.args.a_bytes = (uint32_t)c.a_img.size() / 16 * 16; c.args.b_bytes = (uint32_t)c.b_img.size() / 16 * 16;
  if (c.args.a_bytes > A_REGION || c.args.b_bytes > A_REGION) { printf("%-44s image too large\n", c.name); return 1; }
  c.args.d_cols = c.N;
  k_probe<<<1, 128, SMEM_TOTAL>>>(da, db, dw, dd, c.args);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-44s CUDA ERROR %s\n", c.name, cudaGetErrorString(e)); return 2; }
  std::vector<float> D(128 * c.N);
  cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  const int Kused = c.args.ksteps * 16;
  for (int m = c.row_lo; m < c.row_hi; ++m)
    for (int n = 0; n < c.N; ++n) {
      double r = 0;
      for (int k = 0; k < Kused; ++k) r += (double)c.A[(size_t)m * c.K + k] * c.B[(size_t)n * c.K + k];
      maxerr = fmax(maxerr, fabs(r - D[(size_t)m * c.N + n]));
      maxref = fmax(maxref, fabs(r));
    }
  printf("%-44s max|err| %.3e (max|ref| %.2f)  %s\n", c.name, maxerr, maxref, maxerr < 2e-3 * fmax(maxref, 1.0) ? "OK" : "MISMATCH");
  cudaFree(da); cudaFree(db); cudaFree(dw); cudaFree(dd);
  return maxerr < 2e-3 * fmax(maxref, 1.0) ? 0 : 1;
}

static void fill(Case& c, int M, int N, int K) {
  c.M = M; c.N = N; c.K = K;
  c.A.resize((size_t)M * K); c.B.resize((size_t)N * K);
  for (auto& v : c.A) v = f16(frand());
  for (auto& v : c.B) v = f16(frand());
}

int main(int argc, char** argv) {
  if (argc > 1) g_want = atoi(argv[1]);   // run one case per process: a faulting descriptor poisons the context
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL);
  int bad = 0;
  {  // 1. baseline: K-major SW128 A (128 x 32 of a 64-wide row) and B (144 rows), two k-steps
    Case c; c.name = "kmajor sw128  M128 N144 K32 (+0)"; fill(c, 128, 144, 64);
    for (int m = 0; m < 128; ++m) for (int k = 0; k < 64; ++k) put16(c.a_img, kmaj(m, k, 128), c.A[m * 64 + k]);
    for (int n = 0; n < 144; ++n) for (int k = 0; k < 64; ++k) put16(c.b_img, kmaj(n, k, 144), c.B[n * 64 + k]);
    c.args = ProbeArgs{}; c.args.desc_a_hi = c.args.desc_b_hi = desc_hi(16, 1024, 2);
    c.args.a_inc = c.args.b_inc = 32; c.args.idesc = idesc_f16(128, 144, 0, 0); c.args.ksteps = 2;
    bad += run(c);
    // 1b. second head of the pair: k-steps 2, 3 (start + 64 B)
    Case d = c; d.name = "kmajor sw128  M128 N144 K32 (+64 B: 2nd head)";
    for (int m = 0; m < 128; ++m) for (int k = 0; k < 32; ++k) d.A[m * 64 + k] = c.A[m * 64 + 32 + k];
    for (int n = 0; n < 144; ++n) for (int k = 0; k < 32; ++k) d.B[n * 64 + k] = c.B[n * 64 + 32 + k];
    d.args.a_start = d.args.b_start = 64;
    bad += run(d);
  }
  {  // 2. A tile starting 16 / 112 rows into a 176-row image (leftover-row tiles): rows beyond the image are garbage
    for (int off : {16, 112, 144}) {
      Case c; char* nm = new char[64]; snprintf(nm, 64, "kmajor sw128  A start +%d rows", off); c.name = nm;
      fill(c, 128, 144, 64);
      std::vector<float> Abig((size_t)304 * 64);
      for (auto& v : Abig) v = f16(frand());
      for (int m = 0; m < 304; ++m) for (int k = 0; k < 64; ++k) put16(c.a_img, kmaj(m, k, 304), Abig[m * 64 + k]);
      for (int m = 0; m < 128; ++m) for (int k = 0; k < 64; ++k) c.A[m * 64 + k] = Abig[(m + off) * 64 + k];
      for (int n = 0; n < 144; ++n) for (int k = 0; k < 64; ++k) put16(c.b_img, kmaj(n, k, 144), c.B[n * 64 + k]);
      c.args = ProbeArgs{}; c.args.desc_a_hi = c.args.desc_b_hi = desc_hi(16, 1024, 2);
      c.args.a_start = off * 128; c.args.a_inc = c.args.b_inc = 32; c.args.idesc = idesc_f16(128, 144, 0, 0); c.args.ksteps = 2;
      bad += run(c);
    }
  }
  // 3./4. PV: D[128, N] = P[128, 144] V[144, N];  V image [144 keys][128 B = 64 dims] read as an MN-major B operand
  for (int variant = 0; variant < 6; ++variant) {
    const int N = variant == 2 || variant == 5 ? 64 : 32;
    const int noff = (variant == 1 || variant == 4) ? 32 : 0;       // dims [noff, noff + N) of the 64-wide row
    const bool ts = variant >= 3;
    Case c; char* nm = new char[96];
    snprintf(nm, 96, "PV %s  B mn-major sw128 N%d dims+%d", ts ? "A=TMEM" : "A=smem", N, noff); c.name = nm;
    fill(c, 128, N, 144);
    std::vector<float> V((size_t)144 * 64);
    for (auto& v : V) v = f16(frand());
    for (int k = 0; k < 144; ++k) for (int n = 0; n < 64; ++n) put16(c.b_img, mnmaj(k, n), V[k * 64 + n]);
    for (int n = 0; n < N; ++n) for (int k = 0; k < 144; ++k) c.B[(size_t)n * 144 + k] = V[k * 64 + noff + n];
    c.args = ProbeArgs{};
    c.args.desc_b_hi = desc_hi(16, 1024, 2);
    c.args.b_start = noff * 2; c.args.b_inc = 16 * 128;
    c.args.idesc = idesc_f16(128, N, 0, 1); c.args.ksteps = 9;
    if (!ts) {
      for (int m = 0; m < 128; ++m) for (int k = 0; k < 144; ++k) put16(c.a_img, kmaj(m, k, 128), c.A[m * 144 + k]);
      c.args.desc_a_hi = desc_hi(16, 1024, 2);
      // k-step advance inside a 64-wide k-block is +32 B, from block to block +16 KB: not a constant stride -> probe the
      // first k-block only (4 k-steps) here
      c.args.a_inc = 32; c.args.ksteps = 4;
    } else {
      c.args.a_tmem = 1; c.args.a_cols = 72; c.args.a_inc = 8;
      c.a_words.resize(128 * 72);
      for (int m = 0; m < 128; ++m)
        for (int j = 0; j < 72; ++j)
          c.a_words[m * 72 + j] = (uint32_t)h16(c.A[m * 144 + 2 * j]) | ((uint32_t)h16(c.A[m * 144 + 2 * j + 1]) << 16);
    }
    bad += run(c);
  }
  {  // 5. K-major SWIZZLE_64B (64-byte rows = one head): informational
    Case c; c.name = "kmajor sw64   M128 N144 K32"; fill(c, 128, 144, 32);
    for (int m = 0; m < 128; ++m) for (int k = 0; k < 32; ++k) put16(c.a_img, kmaj64(m, k), c.A[m * 32 + k]);
    for (int n = 0; n < 144; ++n) for (int k = 0; k < 32; ++k) put16(c.b_img, kmaj64(n, k), c.B[n * 32 + k]);
    c.args = ProbeArgs{}; c.args.desc_a_hi = c.args.desc_b_hi = desc_hi(16, 512, 4);
    c.args.a_inc = c.args.b_inc = 32; c.args.idesc = idesc_f16(128, 144, 0, 0); c.args.ksteps = 2;
    run(c);
  }
  {  // 6. accumulate two products with different A/B into one accumulator (leftover rows of two heads)
    Case c; c.name = "two heads into one accumulator (zero rows)"; fill(c, 128, 144, 64);
    // rows 0..15 carry head 0 (k 0..31, zero in k 32..63), rows 16..31 head 1 (k 32..63, zero in k 0..31): accumulating
    // the four k-steps gives each 16-row group its own head's product (the zero rows add nothing)
    for (int n = 0; n < 144; ++n) for (int k = 0; k < 64; ++k) put16(c.b_img, kmaj(n, k, 144), c.B[n * 64 + k]);
    std::vector<float> A2((size_t)128 * 64, 0.f);
    for (int r = 0; r < 16; ++r) for (int k = 0; k < 32; ++k) A2[r * 64 + k] = f16(frand());
    for (int r = 16; r < 32; ++r) for (int k = 32; k < 64; ++k) A2[r * 64 + k] = f16(frand());
    for (int m = 0; m < 128; ++m) for (int k = 0; k < 64; ++k) put16(c.a_img, kmaj(m, k, 128), A2[m * 64 + k]);
    c.A = A2;
    c.args = ProbeArgs{}; c.args.desc_a_hi = c.args.desc_b_hi = desc_hi(16, 1024, 2);
    c.args.a_inc = c.args.b_inc = 32; c.args.idesc = idesc_f16(128, 144, 0, 0); c.args.ksteps = 4;
    c.row_hi = 32;
    bad += run(c);
  }
  // 7. MN-major B wider than one 64-element swizzle atom (SFNO: data as the B operand, N = channels): atoms of
  //    [64 k-rows][128 B] stacked along N, `blk` bytes apart.  Which descriptor field carries the atom stride?
  for (int variant = 0; variant < 4; ++variant) {
    const int N = variant < 2 ? 192 : 256;
    const bool lbo_is_atom_stride = (variant & 1) == 0;
    Case c; char* nm = new char[96];
    snprintf(nm, 96, "B mn-major sw128 N%d, atom stride 8 KB in %s", N, lbo_is_atom_stride ? "LBO (SBO 1024)" : "SBO (LBO 1024)");
    c.name = nm;
    fill(c, 128, N, 64);
    for (int m = 0; m < 128; ++m) for (int k = 0; k < 64; ++k) put16(c.a_img, kmaj(m, k, 128), c.A[m * 64 + k]);
    for (int n = 0; n < N; ++n) for (int k = 0; k < 64; ++k) put16(c.b_img, (size_t)(n / 64) * 8192 + mnmaj(k, n % 64), c.B[(size_t)n * 64 + k]);
    c.args = ProbeArgs{};
    c.args.desc_a_hi = desc_hi(16, 1024, 2);
    c.args.desc_b_hi = lbo_is_atom_stride ? desc_hi(8192, 1024, 2) : desc_hi(1024, 8192, 2);
    c.args.a_inc = 32; c.args.b_inc = 16 * 128;
    c.args.idesc = idesc_f16(128, N, 0, 1); c.args.ksteps = 4;
    bad += run(c);
  }
  printf("probe finished: %d mismatching case(s)\n", bad);
  return 0;
}
