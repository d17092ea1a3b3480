// fp16 GEMM on the 5th-generation tensor cores (tcgen05 + TMEM accumulators), operands fed by TMA.
// Path A of SURVEY.md section 8 (rows A2-A4, A6): every Linear / 1x1 conv / 3x3 conv of the Zero123 UNet and the
// VAE, and the QK^T / PV products of the unfused attention fallback, go through this file.
//
//   C[M,N] = epilogue( alpha * A[M,K] . B[N,K]^T + bias[N] + rowbias[row / rpg, N] ) (+ residual[M,N])   fp16 in, fp32 acc
//
// A and B are both K-major (row-major activations [rows, K]; nn.Linear / flattened conv weights [N, K]).
//
// Main kernel (gemm2): a CTA PAIR (cluster of 2 on one TPC) computes a 256 x BN output tile with
// tcgen05.mma.cta_group::2: each CTA stages its own 128 rows of A and HALF of the B tile, the tensor cores of both SMs
// read both halves, so the L2 -> SM operand traffic per flop is 100-128 flop/B instead of the 64 flop/B of a lone 128 x 128
// tile (r1 ncu: the single-CTA kernel was L2-bandwidth bound at 3-600 TFLOP/s).  Per CTA:
//   warp 0     TMA producer: cp.async.bulk.tensor (cta_group::2 form: completes on the LEADER's mbarrier), SWIZZLE_128B,
//              into a multi-stage shared-memory ring guarded by full (leader) / empty (both CTAs, multicast commit) mbarriers;
//   warp 1     allocates TMEM for the pair; in the leader CTA one lane issues tcgen05.mma (M=256, N=BN, K=16) four
//              times per stage, committing each stage back to both producers and the finished accumulator to both epilogues;
//   warps 2-5  epilogue on the CTA's own 128 accumulator rows: tcgen05.ld (32 lanes x 32 columns) -> bias / row-group bias /
//              activation / GEGLU gate / residual in registers -> fp16 or fp32 stores (rows past M, columns past N masked;
//              TMA zero-fills the K, M and N tails on the way in).
// Two pairs are co-resident per SM pair (<= 104 KB smem, <= 256 TMEM columns each), so one tile's prologue / epilogue
// overlaps the other's main loop.  The single-CTA kernel (gemm1) remains for N <= 64 and for the 4-D batched (per-head) mode.
#include <cuda.h>
#include <stdlib.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace o2345 {
namespace {

constexpr int BM = 128, BK = 64;
constexpr int GEMM_THREADS = 192;
constexpr uint32_t SPIN_LIMIT = 1u << 28;  // bounded waits: a protocol bug traps instead of hanging the GPU
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address: "the even CTA of my pair"

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!done && ++spins > SPIN_LIMIT) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// pair forms: the data lands in the executing CTA's shared memory, the bytes are counted on the leader CTA's barrier
__device__ __forceinline__ void tma2_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4,
// LBO = 1 (ignored for swizzled K-major), SBO = 1024 B (8 rows x 128 B), version 1, layout type 2.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: D = f32, A = B = f16, both K-major, N >> 3 at bit 17, M >> 4 at bit 24.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
  return (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs of the pair once the pair's MMAs so far have finished
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\t"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}"
      ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct GemmParams {
  int M, N, K;
  int64_t ldc;                 // elements between consecutive rows of C / residual
  int nh;                      // batched mode: blockIdx.z = b * nh + h
  int64_t stride_c_h, stride_c_b;  // element offsets of C (and residual) per inner / outer batch index
  const float* bias;           // [N] or nullptr
  const __half* rowbias;       // [groups, rowbias_ld] or nullptr: added before the activation, group = row / rows_per_group
  int64_t rowbias_ld;
  int rows_per_group;
  const __half* residual;      // [M, ldc] or nullptr, added after the activation
  void* C;
  int out_f32;                 // 0: fp16 output, 1: fp32 output
  int act;                     // 0 none, 1 SiLU, 2 GELU(erf), 4 QuickGELU x sigmoid(1.702 x), 3 GEGLU: columns come in chunks of 32 = 16 values + their 16 gates,
                               //    out[:, 16 j + e] = v_e * gelu(g_e); C has N / 2 columns
  float alpha;                 // scale applied to the accumulator before bias
  int batched;                 // 4-D tensor maps (K, rows, h, b)
  // implicit 3x3 convolution (stride 1, zero padding 1): A is the channel-last activation [B, H, W, C] seen through a
  // 4-D tensor map (C, W, H, B); an output tile of 128 consecutive pixels is a box (64 ch, tw, th, tb), and kernel tap
  // (ky, kx) is the same box shifted by (kx-1, ky-1) -- TMA's out-of-bounds zero fill IS the convolution padding.
  int conv, cC, cH, cW, cblocks;
  // split-K for shapes that cannot fill the GPU with output tiles (M <= 2048 with K up to 23 040): `splits` CTAs per
  // tile accumulate with fp32 reductions into ws [M, N] (zero on entry), a second kernel applies the epilogue and re-zeroes.
  int splits;
  float* ws;
  int staged;                  // fp16 output through a shared-memory transpose so that global stores are coalesced rows
  long long* trace;            // diagnostic: CTA (0,0,0) stores clock64() stamps of its phases (o2345_debug_gemm_trace), else nullptr
};

__device__ __forceinline__ void stamp(const GemmParams& p, int slot) {
  if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) p.trace[slot] = clock64();
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == 1) return x / (1.f + __expf(-x));
  if (act == 2) return gelu_erf(x);
  if (act == 4) return x / (1.f + __expf(-1.702f * x));   // QuickGELU (CLIP)
  return x;
}

__device__ __forceinline__ void store8(const GemmParams& p, int64_t off, const float (&v)[8]) {
  if (p.out_f32) {
    float* o = reinterpret_cast<float*>(p.C) + off;
    *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    __half2 h[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.C) + off) = *reinterpret_cast<uint4*>(h);
  }
}

// One thread's 32 consecutive accumulator columns [col0, col0 + 32) of output row `row` (crow = element offset of the row
// in C / residual): everything after the MMA.
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, const uint32_t (&r)[32], int row, int64_t crow, int col0) {
  if (p.splits > 1) {  // partial tile: fp32 reductions into the workspace, epilogue applied by splitk_finalize
    float* w = p.ws + (int64_t)row * p.N + col0;
    if ((p.N & 3) == 0) {
#pragma unroll
      for (int e = 0; e < 32; e += 4)
        if (col0 + e < p.N)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(w + e), "f"(__uint_as_float(r[e])),
                       "f"(__uint_as_float(r[e + 1])), "f"(__uint_as_float(r[e + 2])), "f"(__uint_as_float(r[e + 3]))
                       : "memory");
    } else {
#pragma unroll
      for (int e = 0; e < 32; ++e)
        if (col0 + e < p.N) atomicAdd(w + e, __uint_as_float(r[e]));
    }
    return;
  }
  const __half* rb = p.rowbias ? p.rowbias + (int64_t)(row / p.rows_per_group) * p.rowbias_ld : nullptr;
  if (p.act == 3) {  // GEGLU: 16 values then their 16 gates; N is a multiple of 32 (checked on the host)
    if (col0 >= p.N) return;
    float v[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      float a = __uint_as_float(r[e]) * p.alpha, g = __uint_as_float(r[16 + e]) * p.alpha;
      if (p.bias) a += __ldg(p.bias + col0 + e), g += __ldg(p.bias + col0 + 16 + e);
      v[e] = a * gelu_erf(g);
    }
    const int64_t o = crow + (col0 >> 1);
    float lo[8], hi[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) lo[e] = v[e], hi[e] = v[8 + e];
    store8(p, o, lo);
    store8(p, o + 8, hi);
    return;
  }
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int col = col0 + j;
    if (col >= p.N) break;
    float v[8];
    const bool full = col + 8 <= p.N;
    if (full && rb) {
      uint4 q = *reinterpret_cast<const uint4*>(rb + col);
      const __half* h = reinterpret_cast<const __half*>(&q);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = __half2float(h[e]);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = (rb && col + e < p.N) ? __half2float(rb[col + e]) : 0.f;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float x = fmaf(__uint_as_float(r[j + e]), p.alpha, v[e]);
      if (p.bias && col + e < p.N) x += __ldg(p.bias + col + e);
      v[e] = apply_act(x, p.act);
    }
    if (full && (p.ldc & 7) == 0) {
      if (p.residual) {
        uint4 q = *reinterpret_cast<const uint4*>(p.residual + crow + col);
        const __half* h = reinterpret_cast<const __half*>(&q);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] += __half2float(h[e]);
      }
      store8(p, crow + col, v);
    } else {
      for (int e = 0; e < 8 && col + e < p.N; ++e) {
        float x = v[e];
        if (p.residual) x += __half2float(p.residual[crow + col + e]);
        if (p.out_f32) reinterpret_cast<float*>(p.C)[crow + col + e] = x;
        else reinterpret_cast<__half*>(p.C)[crow + col + e] = __float2half_rn(x);
      }
    }
  }
}

constexpr int EPI_RB_GROUPS = 8;                 // row-bias groups (images) one 128-row tile may span when staged in smem
constexpr int epi_smem_bytes(int bn) { return bn * 4 + EPI_RB_GROUPS * bn * 2; }

// Called by the four epilogue warps while the main loop runs: the tile's bias and row-group-bias slices go to shared
// memory, so that the epilogue proper never waits on a first-touch global load (r1 trace: five serialized L2 misses).
template <int BN>
__device__ __forceinline__ void epilogue_preload(const GemmParams& p, int m0, int n0, float* sbias, __half* srb, int tid_e) {
  if (p.splits <= 1) {
    for (int i = tid_e; i < BN; i += 128) sbias[i] = (p.bias && n0 + i < p.N) ? __ldg(p.bias + n0 + i) : 0.f;
    if (p.rowbias) {
      const int g0 = m0 / p.rows_per_group;
      const int last = (m0 + BM - 1 < p.M ? m0 + BM - 1 : p.M - 1) / p.rows_per_group;
      const int ng = last - g0 + 1;
      if (ng >= 1 && ng <= EPI_RB_GROUPS)
        for (int i = tid_e; i < ng * BN; i += 128) {
          const int gi = i / BN, c = i - gi * BN;
          srb[i] = (n0 + c < p.N) ? p.rowbias[(int64_t)(g0 + gi) * p.rowbias_ld + n0 + c] : __float2half(0.f);
        }
    }
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");   // epilogue warps only
}

template <int ACT>
__device__ __forceinline__ float act_fn(float x) {
  if (ACT == 1) return __fdividef(x, 1.f + __expf(-x));
  if (ACT == 2) return gelu_erf(x);
  if (ACT == 4) return __fdividef(x, 1.f + __expf(-1.702f * x));
  return x;
}

// Phase 1 of the staged epilogue for one thread (= one accumulator row): TMEM -> registers -> alpha / bias / row-group
// bias / activation (ACT 3: GEGLU gate) -> fp16 -> this row of the warp's shared-memory slab.
template <int BN, int ACT>
__device__ __forceinline__ void stage_rows(const GemmParams& p, uint32_t tmem_row_base, uint8_t* mine, int n0, const float* sbias,
                                           const __half* rb, bool rb_smem) {
#pragma unroll 1
  for (int c0 = 0; c0 < BN; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(tmem_row_base + c0, r);
    if (n0 + c0 >= p.N) break;                     // warp-uniform
    if (ACT == 3) {
      __half2 h[8];
#pragma unroll
      for (int e = 0; e < 16; e += 2) {
        float a0 = fmaf(__uint_as_float(r[e]), p.alpha, sbias[c0 + e]), a1 = fmaf(__uint_as_float(r[e + 1]), p.alpha, sbias[c0 + e + 1]);
        float g0 = fmaf(__uint_as_float(r[16 + e]), p.alpha, sbias[c0 + 16 + e]);
        float g1 = fmaf(__uint_as_float(r[17 + e]), p.alpha, sbias[c0 + 17 + e]);
        h[e >> 1] = __floats2half2_rn(a0 * gelu_erf(g0), a1 * gelu_erf(g1));
      }
      uint4* d = reinterpret_cast<uint4*>(mine + (c0 >> 1) * 2);
      d[0] = reinterpret_cast<uint4*>(h)[0], d[1] = reinterpret_cast<uint4*>(h)[1];
    } else {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        float v[8];
        const float4 b0 = *reinterpret_cast<const float4*>(sbias + c0 + j), b1 = *reinterpret_cast<const float4*>(sbias + c0 + j + 4);
        v[0] = b0.x, v[1] = b0.y, v[2] = b0.z, v[3] = b0.w, v[4] = b1.x, v[5] = b1.y, v[6] = b1.z, v[7] = b1.w;
        if (rb && (rb_smem || n0 + c0 + j + 8 <= p.N)) {
          uint4 q = *reinterpret_cast<const uint4*>(rb + c0 + j);
          const __half* hq = reinterpret_cast<const __half*>(&q);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] += __half2float(hq[e]);
        }
        __half2 h[4];
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          float x0 = fmaf(__uint_as_float(r[j + e]), p.alpha, v[e]), x1 = fmaf(__uint_as_float(r[j + e + 1]), p.alpha, v[e + 1]);
          h[e >> 1] = __floats2half2_rn(act_fn<ACT>(x0), act_fn<ACT>(x1));
        }
        *reinterpret_cast<uint4*>(mine + (c0 + j) * 2) = *reinterpret_cast<uint4*>(h);
      }
    }
  }
}

// Staged epilogue of one warp's 32 accumulator rows x BN columns (fp16 output).  Phase 1: thread = row, as the TMEM
// load delivers it: alpha / bias / row-group bias / activation / GEGLU gate, rounded to fp16 (where autocast rounds the
// layer output) into this warp's slab of the (now idle) operand ring.  Phase 2: the warp walks the slab in 16-byte
// pieces along rows, adds the residual and writes whole rows: every global access is a run of full 32-byte sectors, and
// the residual loads of four pieces are in flight together.
// (r1 trace: with one 16-byte store per lane to 32 different rows the epilogue took 46 000 cycles per tile, 10x the
// main loop of a K = 320 GEMM.)
template <int BN>
__device__ __forceinline__ void epilogue_staged(const GemmParams& p, uint32_t tmem_row_base, uint8_t* slab, int lane, int m0,
                                                int row0, int n0, const float* sbias, const __half* srb) {
  const bool geglu = p.act == 3;
  const int outc = geglu ? BN / 2 : BN;           // output columns of this tile
  const int stride = outc * 2 + 16;               // bytes per staged row (+16: 16-byte pieces of consecutive rows rotate banks)
  const int row = row0 + lane;
  // row-group bias of this thread's row: from the smem copy when the tile spans few groups, else straight from global
  const __half* rb = nullptr;
  bool rb_smem = false;
  if (p.rowbias) {
    const int rr = row < p.M ? row : p.M - 1;
    const int g0 = m0 / p.rows_per_group, last = (m0 + BM - 1 < p.M ? m0 + BM - 1 : p.M - 1) / p.rows_per_group;
    rb_smem = last - g0 + 1 <= EPI_RB_GROUPS;
    rb = rb_smem ? srb + (rr / p.rows_per_group - g0) * BN : p.rowbias + (int64_t)(rr / p.rows_per_group) * p.rowbias_ld + n0;
  }
  uint8_t* mine = slab + lane * stride;
  // the activation switch is hoisted out of the element loops: one branch per tile instead of two per element
  switch (p.act) {
    case 1: stage_rows<BN, 1>(p, tmem_row_base, mine, n0, sbias, rb, rb_smem); break;
    case 2: stage_rows<BN, 2>(p, tmem_row_base, mine, n0, sbias, rb, rb_smem); break;
    case 3: stage_rows<BN, 3>(p, tmem_row_base, mine, n0, sbias, rb, rb_smem); break;
    case 4: stage_rows<BN, 4>(p, tmem_row_base, mine, n0, sbias, rb, rb_smem); break;
    default: stage_rows<BN, 0>(p, tmem_row_base, mine, n0, sbias, rb, rb_smem); break;
  }
  __syncwarp();
  const int ppr = outc >> 3;                       // 16-byte pieces per row
  const int total = 32 * ppr;
  const int nout = geglu ? p.N >> 1 : p.N, n0out = geglu ? n0 >> 1 : n0;
  __half* C = reinterpret_cast<__half*>(p.C);
  constexpr int UN = 4;
  for (int base = lane; base < total; base += 32 * UN) {
    uint4 v[UN], q[UN];
    int64_t o[UN];
    bool ok[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int pp = base + 32 * u;
      const int rl = pp / ppr, ci = pp - rl * ppr;
      const int grow = row0 + rl, col = n0out + ci * 8;
      ok[u] = pp < total && grow < p.M && col < nout;   // N is a multiple of 8 on this path (host check): no partial pieces
      o[u] = (int64_t)grow * p.ldc + col;
      if (ok[u]) {
        v[u] = *reinterpret_cast<const uint4*>(slab + rl * stride + ci * 16);
        if (p.residual) q[u] = *reinterpret_cast<const uint4*>(p.residual + o[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      if (!ok[u]) continue;
      if (p.residual) {
        __half2* a = reinterpret_cast<__half2*>(&v[u]);
        const __half2* b = reinterpret_cast<const __half2*>(&q[u]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 fa = __half22float2(a[e]), fb = __half22float2(b[e]);
          a[e] = __floats2half2_rn(fa.x + fb.x, fa.y + fb.y);
        }
      }
      *reinterpret_cast<uint4*>(C + o[u]) = v[u];
    }
  }
}

// ------------------------------------------------------------------------------------------------ single-CTA kernel
template <int BN, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_f16_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by SWIZZLE_128B
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float* sbias = reinterpret_cast<float*>(tmem_slot + 2);
  __half* srb = reinterpret_cast<__half*>(sbias + BN);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN, bz = blockIdx.z;
  const int nk = p.conv ? 9 * p.cblocks : (p.K + BK - 1) / BK;
  // split-K: blockIdx.z owns k-blocks [kb0, kb1) and adds its partial tile into the fp32 workspace
  int kb0 = 0, kb1 = nk;
  if (p.splits > 1) {
    kb0 = (int)((int64_t)nk * bz / p.splits);
    kb1 = (int)((int64_t)nk * (bz + 1) / p.splits);
  }

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < STAGES; ++s) mbar_init(full + s, 1), mbar_init(empty + s, 1);
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: BN fp32 columns x 128 lanes
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // everything above touched no global memory: it overlaps the tail of the previous kernel
  pdl_trigger();

  if (warp == 0) {
    if (lane == 0) {  // ---------------- TMA producer
      for (int kb = kb0; kb < kb1; ++kb) {
        int s = (kb - kb0) % STAGES;
        uint32_t ph = ((kb - kb0) / STAGES) & 1;
        mbar_wait(empty + s, ph ^ 1);
        mbar_expect_tx(full + s, A_BYTES + B_BYTES);
        if (p.conv) {
          const int tap = kb / p.cblocks, c0 = (kb - tap * p.cblocks) * BK;
          const int x0 = m0 % p.cW, y0 = (m0 / p.cW) % p.cH, b0 = m0 / (p.cW * p.cH);
          tma_load_4d(sA + s * A_BYTES, &tmA, full + s, c0, x0 + tap % 3 - 1, y0 + tap / 3 - 1, b0);
          tma_load_2d(sB + s * B_BYTES, &tmB, full + s, tap * p.cC + c0, n0);
        } else if (p.batched) {
          tma_load_4d(sA + s * A_BYTES, &tmA, full + s, kb * BK, m0, bz % p.nh, bz / p.nh);
          tma_load_4d(sB + s * B_BYTES, &tmB, full + s, kb * BK, n0, bz % p.nh, bz / p.nh);
        } else {
          tma_load_2d(sA + s * A_BYTES, &tmA, full + s, kb * BK, m0);
          tma_load_2d(sB + s * B_BYTES, &tmB, full + s, kb * BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---------------- MMA issuer
      constexpr uint32_t idesc = umma_idesc_f16(BM, BN);
      for (int kb = kb0; kb < kb1; ++kb) {
        int s = (kb - kb0) % STAGES;
        uint32_t ph = ((kb - kb0) / STAGES) & 1;
        mbar_wait(full + s, ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t a0 = smem_u32(sA + s * A_BYTES), b0 = smem_u32(sB + s * B_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // advancing 16 fp16 along K inside the 128-byte swizzle atom = +32 bytes on the start address
          umma_f16(tmem_base, umma_desc_sw128(a0 + k * 32), umma_desc_sw128(b0 + k * 32), idesc, ((kb - kb0) | k) != 0);
        }
        umma_commit(empty + s);   // frees the smem stage once these MMAs have read it
      }
      umma_commit(tmem_full);     // accumulator complete
    }
  } else {  // ------------------------ epilogue warps 2..5, TMEM lane quarter = warp % 4
    const int quarter = warp & 3;
    if (p.staged) epilogue_preload<BN>(p, m0, n0, sbias, srb, threadIdx.x - 64);
    mbar_wait(tmem_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = m0 + quarter * 32 + lane;
    if (p.staged) {   // operand ring is idle once the accumulator is complete: reuse it for the output transpose
      epilogue_staged<BN>(p, tmem_base + ((uint32_t)(quarter * 32) << 16), smem + quarter * 32 * (BN * 2 + 16), lane, m0,
                          m0 + quarter * 32, n0, sbias, srb);
    } else {
      const int64_t crow = (p.batched ? (int64_t)(bz % p.nh) * p.stride_c_h + (int64_t)(bz / p.nh) * p.stride_c_b : 0) +
                           (int64_t)row * p.ldc;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + c0, r);
        if (row < p.M) epilogue_chunk(p, r, row, crow, n0 + c0);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BN));
  }
}

// ------------------------------------------------------------------------------------------------ CTA-pair kernel
__host__ __device__ constexpr int tmem_cols(int bn) { return bn <= 32 ? 32 : bn <= 64 ? 64 : bn <= 128 ? 128 : bn <= 256 ? 256 : 512; }

template <int BN, int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 2)
gemm2_f16_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int BH = BN / 2;                                  // rows of B staged by each CTA of the pair
  constexpr int A_BYTES = BM * BK * 2, B_BYTES = BH * BK * 2;
  static_assert(B_BYTES % 1024 == 0, "stage bases must stay 1024-byte aligned for SWIZZLE_128B");
  constexpr int TCOLS = tmem_cols(BN);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float* sbias = reinterpret_cast<float*>(tmem_slot + 2);
  __half* srb = reinterpret_cast<__half*>(sbias + BN);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) stamp(p, 0);
  const uint32_t rank = cluster_ctarank();                    // 0 = leader (issues the MMAs, owns the full barriers)
  const int m0 = (blockIdx.x >> 1) * (2 * BM) + (int)rank * BM, n0 = blockIdx.y * BN, bz = blockIdx.z;
  const int nk = p.conv ? 9 * p.cblocks : (p.K + BK - 1) / BK;
  int kb0 = 0, kb1 = nk;
  if (p.splits > 1) {
    kb0 = (int)((int64_t)nk * bz / p.splits);
    kb1 = (int)((int64_t)nk * (bz + 1) / p.splits);
  }

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < STAGES; ++s) mbar_init(full + s, 1), mbar_init(empty + s, 1);
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {  // the same warp of BOTH CTAs allocates the pair's TMEM columns
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TCOLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();   // barrier inits of the leader must be visible before the peer's TMA can complete on them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // everything above touched no global memory: it overlaps the tail of the previous kernel
  pdl_trigger();
  if (threadIdx.x == 0) stamp(p, 1);

  if (warp == 0) {
    if (lane == 0) {  // ---------------- TMA producer (both CTAs): own 128 rows of A, own half of the B tile
      for (int kb = kb0; kb < kb1; ++kb) {
        int s = (kb - kb0) % STAGES;
        uint32_t ph = ((kb - kb0) / STAGES) & 1;
        mbar_wait(empty + s, ph ^ 1);
        if (rank == 0) mbar_expect_tx(full + s, 2 * (A_BYTES + B_BYTES));   // the peer's bytes land on this barrier too
        const int nb = n0 + (int)rank * BH;
        if (p.conv) {
          const int tap = kb / p.cblocks, c0 = (kb - tap * p.cblocks) * BK;
          const int x0 = m0 % p.cW, y0 = (m0 / p.cW) % p.cH, b0 = m0 / (p.cW * p.cH);
          tma2_load_4d(sA + s * A_BYTES, &tmA, full + s, c0, x0 + tap % 3 - 1, y0 + tap / 3 - 1, b0);
          tma2_load_2d(sB + s * B_BYTES, &tmB, full + s, tap * p.cC + c0, nb);
        } else {
          tma2_load_2d(sA + s * A_BYTES, &tmA, full + s, kb * BK, m0);
          tma2_load_2d(sB + s * B_BYTES, &tmB, full + s, kb * BK, nb);
        }
        if (kb == kb0) stamp(p, 2);
      }
      stamp(p, 3);
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {  // ---------------- MMA issuer (leader only): M = 256 across the pair
      constexpr uint32_t idesc = umma_idesc_f16(2 * BM, BN);
      for (int kb = kb0; kb < kb1; ++kb) {
        int s = (kb - kb0) % STAGES;
        uint32_t ph = ((kb - kb0) / STAGES) & 1;
        mbar_wait(full + s, ph);
        if (kb == kb0) stamp(p, 4);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t a0 = smem_u32(sA + s * A_BYTES), b0 = smem_u32(sB + s * B_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          umma2_f16(tmem_base, umma_desc_sw128(a0 + k * 32), umma_desc_sw128(b0 + k * 32), idesc, ((kb - kb0) | k) != 0);
        umma2_commit(empty + s);   // frees this stage in both CTAs
      }
      umma2_commit(tmem_full);     // accumulator complete: both epilogues may start
      stamp(p, 5);
    }
  } else {  // ------------------------ epilogue warps 2..5 on this CTA's 128 accumulator rows
    const int quarter = warp & 3;
    if (p.staged) epilogue_preload<BN>(p, m0, n0, sbias, srb, threadIdx.x - 64);
    mbar_wait(tmem_full, 0);
    if (threadIdx.x == 64) stamp(p, 6);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = m0 + quarter * 32 + lane;
    if (p.staged) {   // operand ring is idle once the accumulator is complete: reuse it for the output transpose
      epilogue_staged<BN>(p, tmem_base + ((uint32_t)(quarter * 32) << 16), smem + quarter * 32 * (BN * 2 + 16), lane, m0,
                          m0 + quarter * 32, n0, sbias, srb);
    } else {
      const int64_t crow = (int64_t)row * p.ldc;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + c0, r);
        if (row < p.M) epilogue_chunk(p, r, row, crow, n0 + c0);
      }
    }
    if (threadIdx.x == 64) stamp(p, 7);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();   // neither CTA may free TMEM / exit while the pair's MMAs or the peer's TMEM reads are in flight
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TCOLS));
  }
  if (threadIdx.x == 0) stamp(p, 8);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// [nb, nh, rows, K] fp16 view with row stride ld and batch strides sh / sb (elements); box = 64 x box_rows (x 1 x 1)
int make_map(CUtensorMap* m, const void* ptr, int64_t rows, int64_t K, int64_t ld, int nh, int nb, int64_t sh, int64_t sb,
             int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return O2345_ECUDA; }
  cuuint64_t dims[4] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)(nh > 0 ? nh : 1), (cuuint64_t)(nb > 0 ? nb : 1)};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)sh * 2, (cuuint64_t)sb * 2};
  cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  int rank = nh > 0 ? 4 : 2;
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with %d (rows=%lld K=%lld ld=%lld)", (int)r, (long long)rows, (long long)K, (long long)ld); return O2345_ECUDA; }
  return O2345_OK;
}

// split-K epilogue: out = act(alpha * ws + bias + rowbias) + residual, and ws is left zeroed for the next call.
// One thread per (row, 8 columns) when N and ldc allow 16-byte accesses, else one per element.
__global__ void splitk_finalize_kernel(GemmParams p, int vec) {
  pdl_wait();
  pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (vec) {
    const int n8 = p.N >> 3;
    if (i >= p.M * n8) return;
    const int row = i / n8, col = (i - row * n8) << 3;
    float4* w = reinterpret_cast<float4*>(p.ws + (int64_t)row * p.N + col);
    float4 a = w[0], b = w[1];
    w[0] = make_float4(0.f, 0.f, 0.f, 0.f), w[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    const int64_t o = (int64_t)row * p.ldc + col;
    if (p.rowbias) {
      uint4 q = *reinterpret_cast<const uint4*>(p.rowbias + (int64_t)(row / p.rows_per_group) * p.rowbias_ld + col);
      const __half* h = reinterpret_cast<const __half*>(&q);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = fmaf(v[e], p.alpha, __half2float(h[e]));
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] *= p.alpha;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (p.bias) v[e] += __ldg(p.bias + col + e);
      v[e] = apply_act(v[e], p.act);
    }
    if (p.residual) {
      uint4 q = *reinterpret_cast<const uint4*>(p.residual + o);
      const __half* h = reinterpret_cast<const __half*>(&q);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += __half2float(h[e]);
    }
    store8(p, o, v);
    return;
  }
  if (i >= p.M * p.N) return;
  const int row = i / p.N, col = i - row * p.N;
  float x = p.ws[i] * p.alpha;
  p.ws[i] = 0.f;
  if (p.rowbias) x += __half2float(p.rowbias[(int64_t)(row / p.rows_per_group) * p.rowbias_ld + col]);
  if (p.bias) x += __ldg(p.bias + col);
  x = apply_act(x, p.act);
  int64_t o = (int64_t)row * p.ldc + col;
  if (p.residual) x += __half2float(p.residual[o]);
  if (p.out_f32) reinterpret_cast<float*>(p.C)[o] = x;
  else reinterpret_cast<__half*>(p.C)[o] = __float2half_rn(x);
}

// how many k-splits for a non-batched problem whose output tiles give `ctas` CTAs: fill ~2 CTAs per SM, keep >= 4 k-blocks
// per split
int pick_splits(const GemmParams& p, int ctas, float* ws, int64_t ws_floats) {
  static int min_kb = -1;
  if (min_kb < 0) {   // k-blocks each split must keep (tuning knob; the default is the measured optimum of the UNet pass)
    const char* e = getenv("O2345_SPLITK_MIN_KB");
    min_kb = e ? atoi(e) : 12;   // r1 sweep of the UNet pass: 4 -> 5.20 ms, 8 -> 5.03, 12 -> 4.98, 16 -> 5.01, 24 -> 5.11
    if (min_kb < 1) min_kb = 1;
  }
  if (!ws || p.batched || p.act == 3 || (int64_t)p.M * p.N > ws_floats) return 1;
  int nk = p.conv ? 9 * p.cblocks : cdiv(p.K, BK);
  if (ctas >= 120 || nk < 2 * min_kb) return 1;
  int s = 2 * sm_count() / ctas;
  if (s > nk / min_kb) s = nk / min_kb;
  if (s > 32) s = 32;
  return s < 2 ? 1 : s;
}

// coalesced (shared-memory staged) epilogue: fp16 output, whole 16-byte pieces, one pass (no split-K, no head batches)
void pick_staged(GemmParams& p) {
  const int nout = p.act == 3 ? p.N / 2 : p.N;
  p.staged = !p.out_f32 && p.splits <= 1 && !p.batched && (nout % 8) == 0 && (p.ldc % 8) == 0 && ((uintptr_t)p.C % 16) == 0 &&
             (!p.residual || ((uintptr_t)p.residual % 16) == 0);
}

int finalize(const GemmParams& p, cudaStream_t st) {
  if (p.splits <= 1) return O2345_OK;
  const int vec = (p.N % 8) == 0 && (p.ldc % 8) == 0 && (!p.rowbias || (p.rowbias_ld % 8) == 0);
  const int64_t n = vec ? (int64_t)p.M * (p.N / 8) : (int64_t)p.M * p.N;
  O2345_CUDA(launch_pdl(splitk_finalize_kernel, dim3(cdiv(n, 256)), dim3(256), (size_t)(0), st, p, vec));
  O2345_LAUNCH_CHECK();
  return O2345_OK;
}

template <int BN, int STAGES>
int launch1(const CUtensorMap& a, const CUtensorMap& b, const GemmParams& p, int batch, cudaStream_t st) {
  constexpr int SMEM = STAGES * (BM * BK * 2 + BN * BK * 2) + (2 * STAGES + 1) * 8 + 16 + epi_smem_bytes(BN) + 1024;
  static PerDeviceOnce attr;
  if (attr.need()) {
    O2345_CUDA(cudaFuncSetAttribute(gemm_f16_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  }
  dim3 grid(cdiv(p.N, BN), cdiv(p.M, BM), batch > 0 ? batch : (p.splits > 1 ? p.splits : 1));
  O2345_CUDA(launch_pdl(gemm_f16_tc_kernel<BN, STAGES>, dim3(grid), dim3(GEMM_THREADS), (size_t)(SMEM), st, a, b, p));
  O2345_LAUNCH_CHECK();
  return finalize(p, st);
}

template <int BN, int STAGES>
int launch2(const CUtensorMap& a, const CUtensorMap& b, const GemmParams& p, cudaStream_t st) {
  constexpr int SMEM = STAGES * (BM * BK * 2 + (BN / 2) * BK * 2) + (2 * STAGES + 1) * 8 + 16 + epi_smem_bytes(BN) + 1024;
  static PerDeviceOnce attr;
  if (attr.need()) {
    O2345_CUDA(cudaFuncSetAttribute(gemm2_f16_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  }
  dim3 grid(2 * cdiv(p.M, 2 * BM), cdiv(p.N, BN), p.splits > 1 ? p.splits : 1);
  O2345_CUDA(launch_pdl(gemm2_f16_tc_kernel<BN, STAGES>, dim3(grid), dim3(GEMM_THREADS), (size_t)(SMEM), st, a, b, p));
  O2345_LAUNCH_CHECK();
  return finalize(p, st);
}

// tile width of the CTA-pair kernel for an N-column problem: 160 divides every UNet width (320 k), else 256 / 128
int pick_bn2(int N) {
  static int n256 = -1;
  if (n256 < 0) {   // widths from which a multiple of 256 takes the 256-wide tile instead of 160 (tuning knob)
    const char* e = getenv("O2345_BN256_MIN_N");
    n256 = e ? atoi(e) : (1 << 30);   // r1 sweep of the UNet pass: off 4.99 ms, N >= 2560 -> 5.06, N >= 1280 -> 5.23: 160 stays
  }
  if (N <= 64) return 0;          // single-CTA kernel with BN = 64
  if (N % 256 == 0 && N >= n256) return 256;
  if (N % 160 == 0) return 160;
  if (N <= 128) return 128;
  return N % 256 == 0 || N > 640 ? 256 : (N % 128 == 0 ? 128 : 160);
}

long long* g_trace = nullptr;

int fill_epilogue(GemmParams& p, const o2345_epilogue* ep, int M, int N, int64_t ldc) {
  p.bias = nullptr, p.rowbias = nullptr, p.rowbias_ld = 0, p.rows_per_group = 1, p.residual = nullptr;
  p.out_f32 = 0, p.act = 0, p.alpha = 1.f, p.trace = g_trace;
  if (!ep) return O2345_OK;
  O2345_CHECK_ARG(ep->act >= 0 && ep->act <= 4, "unknown activation");
  O2345_CHECK_ARG(!ep->rowbias || (ep->rows_per_group > 0 && (ep->rowbias_ld % 8) == 0 && ((uintptr_t)ep->rowbias % 16) == 0),
                  "row bias: rows_per_group > 0, 16-byte aligned, row stride a multiple of 8");
  O2345_CHECK_ARG(ep->act != 3 || ((N % 32) == 0 && (ldc % 8) == 0 && !ep->residual && !ep->rowbias),
                  "GEGLU epilogue: N must be a multiple of 32, ldc of 8, no residual / row bias");
  p.bias = ep->bias, p.rowbias = reinterpret_cast<const __half*>(ep->rowbias), p.rowbias_ld = ep->rowbias_ld;
  p.rows_per_group = ep->rowbias ? ep->rows_per_group : 1;
  p.residual = reinterpret_cast<const __half*>(ep->residual), p.out_f32 = ep->out_f32, p.act = ep->act, p.alpha = ep->alpha;
  (void)M;
  return O2345_OK;
}

int dispatch2(int bn, const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p, cudaStream_t st) {
  if (bn == 160) return launch2<160, 4>(ma, mb, p, st);
  if (bn == 256) return launch2<256, 3>(ma, mb, p, st);
  return launch2<128, 4>(ma, mb, p, st);
}

}  // namespace
}  // namespace o2345

using namespace o2345;

extern "C" void o2345_debug_gemm_trace(long long* device_buf16) { g_trace = device_buf16; }

extern "C" int o2345_conv3x3_f16(const void* x, int B, int H, int W, int C, const void* weight, int N, void* out, int64_t ldc,
                                 const o2345_epilogue* ep, float* splitk_ws, int64_t ws_floats, o2345_stream_t stream) {
  O2345_CHECK_ARG(x && weight && out, "null pointer");
  O2345_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && (C % 8) == 0 && N > 0, "bad sizes (C must be a multiple of 8)");
  // an output tile is 128 consecutive pixels fetched as ONE box (tw, th, tb): either whole multiples of 128 along a row,
  // or whole rows that tile the image exactly (H a multiple of 128 / W), or whole images (H * W divides 128).  Any other
  // shape would wrap a tile across the image border (silently wrong rows) or give a box of fewer than 128 rows (the
  // stage's byte count would never be reached): refused here, callers take the im2col route.
  O2345_CHECK_ARG((W % 128) == 0 || ((128 % W) == 0 && (((int64_t)H * W >= 128 && (H % (128 / W)) == 0) ||
                                                        ((int64_t)H * W < 128 && (128 % (H * W)) == 0))),
                  "implicit 3x3 conv: the image must tile into 128-pixel boxes (W % 128 == 0, or 128 % W == 0 with "
                  "H % (128 / W) == 0, or 128 % (H * W) == 0)");
  O2345_CHECK_ARG(((uintptr_t)x % 16) == 0 && ((uintptr_t)weight % 16) == 0, "operands must be 16-byte aligned");
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return O2345_ECUDA; }
  const int tw = W >= 128 ? 128 : W;
  const int th = W >= 128 ? 1 : (128 / W < H ? 128 / W : H);
  const int tb = 128 / (tw * th);
  CUtensorMap ma, mb;
  {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)tb};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (conv activation) failed with %d", (int)r); return O2345_ECUDA; }
  }
  GemmParams p;
  p.M = B * H * W, p.N = N, p.K = 9 * C, p.ldc = ldc, p.nh = 1, p.stride_c_h = 0, p.stride_c_b = 0, p.C = out;
  int rc = fill_epilogue(p, ep, p.M, N, ldc);
  if (rc) return rc;
  p.batched = 0, p.conv = 1, p.cC = C, p.cH = H, p.cW = W, p.cblocks = (C + BK - 1) / BK;
  p.ws = splitk_ws;
  cudaStream_t st = (cudaStream_t)stream;
  const int bn = pick_bn2(N);
  if (bn == 0) {
    rc = make_map(&mb, weight, N, 9 * (int64_t)C, 9 * (int64_t)C, 0, 0, 0, 0, 64);
    if (rc) return rc;
    p.splits = pick_splits(p, cdiv(p.M, BM), splitk_ws, ws_floats);
    pick_staged(p);
    return launch1<64, 4>(ma, mb, p, 0, st);
  }
  rc = make_map(&mb, weight, N, 9 * (int64_t)C, 9 * (int64_t)C, 0, 0, 0, 0, bn / 2);
  if (rc) return rc;
  p.splits = pick_splits(p, 2 * cdiv(p.M, 2 * BM) * cdiv(N, bn), splitk_ws, ws_floats);
  pick_staged(p);
  return dispatch2(bn, ma, mb, p, st);
}

extern "C" int o2345_gemm_f16(const void* A, const void* B, void* C, int M, int N, int K, int64_t lda, int64_t ldb,
                              int64_t ldc, int nh, int nb, int64_t stride_a_h, int64_t stride_a_b, int64_t stride_b_h,
                              int64_t stride_b_b, int64_t stride_c_h, int64_t stride_c_b, const o2345_epilogue* ep,
                              float* splitk_ws, int64_t ws_floats, o2345_stream_t stream) {
  O2345_CHECK_ARG(A && B && C, "null pointer");
  O2345_CHECK_ARG(M > 0 && N > 0 && K > 0 && nh >= 0 && (nh == 0 || nb >= 1), "bad sizes");
  O2345_CHECK_ARG((lda % 8) == 0 && (ldb % 8) == 0, "row strides of A and B must be multiples of 8 fp16 (16 bytes) for TMA");
  O2345_CHECK_ARG(((uintptr_t)A % 16) == 0 && ((uintptr_t)B % 16) == 0, "A and B must be 16-byte aligned");
  O2345_CHECK_ARG(nh == 0 || ((stride_a_h % 8) == 0 && (stride_a_b % 8) == 0 && (stride_b_h % 8) == 0 && (stride_b_b % 8) == 0),
                  "batch strides must be multiples of 8 fp16");
  GemmParams p;
  p.M = M, p.N = N, p.K = K, p.ldc = ldc, p.nh = nh > 0 ? nh : 1, p.stride_c_h = stride_c_h, p.stride_c_b = stride_c_b, p.C = C;
  int rc = fill_epilogue(p, ep, M, N, ldc);
  if (rc) return rc;
  O2345_CHECK_ARG(nh == 0 || (!p.rowbias && p.act != 3), "row bias / GEGLU are not available in batched mode");
  p.batched = nh > 0 ? 1 : 0;
  p.conv = 0, p.cC = p.cH = p.cW = p.cblocks = 0;
  p.ws = splitk_ws;
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap ma, mb;
  rc = make_map(&ma, A, M, K, lda, nh, nb, stride_a_h, stride_a_b, BM);
  if (rc) return rc;
  const int bn = nh > 0 ? 0 : pick_bn2(N);
  if (bn == 0) {  // single-CTA kernel: per-head batches and N <= 64
    const int BN1 = N <= 64 ? 64 : 128;
    rc = make_map(&mb, B, N, K, ldb, nh, nb, stride_b_h, stride_b_b, BN1);
    if (rc) return rc;
    p.splits = pick_splits(p, cdiv(N, BN1) * cdiv(M, BM), splitk_ws, ws_floats);
    pick_staged(p);
    const int batch = nh > 0 ? nh * nb : 0;
    if (BN1 == 64) return launch1<64, 4>(ma, mb, p, batch, st);
    return launch1<128, 3>(ma, mb, p, batch, st);
  }
  rc = make_map(&mb, B, N, K, ldb, 0, 0, 0, 0, bn / 2);
  if (rc) return rc;
  p.splits = pick_splits(p, 2 * cdiv(M, 2 * BM) * cdiv(N, bn), splitk_ws, ws_floats);
  pick_staged(p);
  return dispatch2(bn, ma, mb, p, st);
}
