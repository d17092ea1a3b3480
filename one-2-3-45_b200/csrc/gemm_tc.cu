// fp16 GEMM on the 5th-generation tensor cores (tcgen05 + TMEM accumulators), operands fed by TMA.
// Path A of SURVEY.md section 8 (rows A2-A4, A6, A8): every Linear / 1x1 conv / 3x3 conv of the Zero123 UNet, the VAE and
// the CLIP tower, and the QK^T / PV products of the unfused attention fallback, go through this file.
//
//   C[M,N] = epilogue( alpha * A[M,K] . B[N,K]^T + bias[N] + rowbias[row / rpg, N] ) (+ residual[M,N])   fp16 in, fp32 acc
//
// A and B are both K-major (row-major activations [rows, K]; nn.Linear / flattened conv weights [N, K]).
//
// ONE kernel template, gemm_tc_kernel<BN, STAGES, CTAS, MODE> (round 1 had two hand-copied kernels):
//   CTAS = 2  a CTA PAIR (cluster of 2 on one TPC) computes a 256 x BN tile with tcgen05.mma.cta_group::2: each CTA stages
//             its own 128 rows of A and HALF of the B tile, the tensor cores of both SMs read both halves (100-128 flop per
//             operand byte instead of 64 for a lone 128 x 128 tile; the L2 -> SM fabric is the binding resource);
//   CTAS = 1  a single CTA computes 128 x BN (M <= 128, and the 4-D batched per-head mode).
// Warp roles (320 threads, two CTAs resident per SM so that one tile's prologue / epilogue overlaps the other's main loop):
//   warp 0      TMA producer: cp.async.bulk.tensor (SWIZZLE_128B) into a STAGES-deep shared-memory ring guarded by full /
//               empty mbarriers (pair: the bytes of both CTAs are counted on the leader's full barrier, stages are freed in
//               both CTAs by a multicast tcgen05.commit);
//   warp 1      allocates TMEM; one lane (of the leader) issues tcgen05.mma M = 128 * CTAS, N = BN, K = 16, four per stage;
//   warps 2-9   EIGHT epilogue warps (round 1: four): warp w owns TMEM lane quarter w % 4 (a hardware rule) and one of two
//               column ranges of the tile.  While the main loop runs they stage bias / row-group bias in shared memory and
//               PREFETCH THE RESIDUAL into registers (round 1 fetched it after the accumulator was complete: 2 400 - 6 000
//               cycles of exposed latency per tile); then tcgen05.ld -> alpha / bias / activation / GEGLU gate -> fp16 ->
//               a shared-memory transpose in the idle operand ring -> coalesced row stores with the residual added.
// MODE selects the epilogue at compile time (round 1 inlined every variant into one 9 600-instruction body):
//   0 staged, no activation   1 staged, GEGLU gate   2 staged, SiLU / GELU / QuickGELU (runtime switch per tile)
//   3 generic (fp32 output, batched, unaligned N) and SPLIT-K.
// Split-K (tiles alone cannot fill 148 SMs): the `splits` CTAs (pairs) of a tile are launched as ONE thread-block cluster
// (CTAS, 1, splits), so the hardware co-schedules them.  Each stores its partial accumulator into its own fp32 plane of the
// workspace, a cluster barrier (release / acquire) publishes the planes, and every split then sums the planes and applies
// the epilogue to ITS share of the tile.  No atomics, no tickets, no zero-initialised scratch, no second kernel (round 1
// added into one plane with L2 atomics -- they sustain only ~90 G elements/s -- and launched a finalize kernel: 92 extra
// launches per UNet iteration; a first round-2 version let the last-arriving CTA finish the tile alone: 13-19 us of
// serial L2 round trips, see profiles/r2_gemm_trace.txt).
// Every mbarrier wait is bounded in TIME (4 s): a protocol bug or a lost arrival records which barrier of which CTA of
// which problem stalled in a host-visible buffer (o2345_last_trap) and traps, instead of spinning for tens of minutes.
#include <cuda.h>
#include <stdlib.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace o2345 {
namespace {

constexpr int BM = 128, BK = 64;
constexpr int EPI_WARPS = 8;
constexpr int EPI_THREADS = 32 * EPI_WARPS;
constexpr int GEMM_THREADS = 64 + EPI_THREADS;
constexpr uint64_t WAIT_LIMIT_NS = 4000000000ull;   // bounded waits: a protocol bug traps (with a record) instead of hanging the GPU
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;         // clears the CTA-rank bit of a shared::cluster address: "the even CTA of my pair"
constexpr int MAX_CLUSTER = 16;                     // CTAS * splits: one cluster per tile (non-portable size, opted into per kernel)
constexpr int RES_PREFETCH = 8;                     // 16-byte residual pieces per lane fetched before the accumulator is ready

enum { WAIT_EMPTY = 1, WAIT_FULL = 2, WAIT_ACC = 3, WAIT_ACC_FREE = 4 };   // which wait timed out (o2345_last_trap)

struct TrapRecord {
  unsigned long long magic;
  int tag, stage, bx, by, bz, rank, M, N, K, bn, ctas, mode, splits, conv;
};
constexpr unsigned long long TRAP_MAGIC = 0x6f32333435545250ull;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done;
}
__device__ __forceinline__ uint64_t global_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
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
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // the eight epilogue warps only

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
template <int CTAS>
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (CTAS == 2)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// pair: arrives on the barrier at this shared-memory offset in BOTH CTAs of the pair (cluster ranks 2j and 2j + 1: `mask`)
// once the pair's MMAs so far have finished
template <int CTAS>
__device__ __forceinline__ void umma_commit(uint64_t* bar, uint16_t mask) {
  if (CTAS == 2)
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
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
  // tap geometry: ctaps taps in rows of ctx, tap j reads the box shifted by (j % ctx + cox, j / ctx + coy).  3 x 3 / pad 1: ctaps 9,
  // ctx 3, cox = coy = -1.  Nearest-neighbour 2x up-sampling followed by a 3 x 3 convolution is FOUR 2 x 2 convolutions of the
  // low-resolution input, one per output phase (a, b) = (row parity, column parity): ctaps 4, ctx 2, cox = b - 1, coy = a - 1,
  // weights pre-summed on the host (rows that collapse onto the same input pixel), and the tile's 128 low-resolution pixels
  // are written to output pixels (2 y + a, 2 x + b): up = 1, upa = a, upb = b.  No im2col buffer, 4 C instead of 9 C per output.
  int ctaps, ctx, cox, coy, up, upa, upb;
  // split-K: the splits of a tile form one cluster; CTA z stores its partial accumulator in plane z of ws [splits, M, N],
  // the cluster barrier publishes the planes, then every split sums and finishes its share of the tile.
  int splits;
  float* ws;
  // GroupNorm statistics of the OUTPUT: colstats[(g * 2 + 0) * N + col] += x and colstats[(g * 2 + 1) * N + col] += x^2 over the
  // final fp16 values of row group g = row / stats_rpg (the consumer turns them into mean / rstd); nullptr: off
  float* colstats;
  int stats_rpg, stats_groups;
  long long* trace;            // diagnostic: CTA (0,0,0) stores clock64() stamps of its phases (o2345_debug_gemm_trace), else nullptr
  TrapRecord* diag;            // host-mapped record written before a bounded wait traps (may be nullptr)
  int bn, ctas, mode;          // for the trap record
};

__device__ __forceinline__ void stamp(const GemmParams& p, int slot) {
  if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) p.trace[slot] = clock64();
}
// wall-clock (globaltimer, ns) stamps of tile (0,0): comparable across the SMs the splits of a tile run on
__device__ __forceinline__ void stamp_ns(const GemmParams& p, int slot, bool any_z = false) {
  if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && (any_z || blockIdx.z == 0)) p.trace[slot] = (long long)global_ns();
}

// row of C that accumulator row `row` is written to (up-sampling phases scatter the low-resolution pixels over the 2x grid)
__device__ __forceinline__ int64_t out_row(const GemmParams& p, int row) {
  if (!p.up) return row;
  const int x = row % p.cW, t = row / p.cW, y = t % p.cH, b = t / p.cH;
  return ((int64_t)b * (2 * p.cH) + 2 * y + p.upa) * (2 * p.cW) + 2 * x + p.upb;
}

__device__ __noinline__ void wait_timed_out(const GemmParams& p, int tag, int stage) {
  if (p.diag) {
    TrapRecord* d = p.diag;
    d->tag = tag, d->stage = stage, d->bx = blockIdx.x, d->by = blockIdx.y, d->bz = blockIdx.z, d->rank = (int)cluster_ctarank();
    d->M = p.M, d->N = p.N, d->K = p.K, d->bn = p.bn, d->ctas = p.ctas, d->mode = p.mode, d->splits = p.splits, d->conv = p.conv;
    __threadfence_system();
    d->magic = TRAP_MAGIC;
    __threadfence_system();
  }
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, const GemmParams& p, int tag, int stage) {
  if (mbar_try(bar, parity)) return;
  const uint64_t t0 = global_ns();
  uint32_t spins = 0;
  while (!mbar_try(bar, parity))
    if (((++spins) & 255u) == 0 && global_ns() - t0 > WAIT_LIMIT_NS) wait_timed_out(p, tag, stage);
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
// GELU(erf) for the GEGLU gate of the staged epilogue, whose result is rounded to fp16 (2^-11 relative) right away: erf by
// Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7 absolute, ~3e-7 with ex2.approx), 14 instructions instead of erff's ~30 with
// range branches -- the GEGLU projections (8192 x 2560 x 320 ...) are bound by the epilogue's ALU work, not by the MMAs.
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, z, 1.f));
  const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-z * z * 1.4426950408889634f));
  const float erf_abs = fmaf(-poly, e, 1.f);
  return 0.5f * x * (1.f + copysignf(erf_abs, x));
}
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

// Generic epilogue (MODE 3, one pass): one thread's 32 consecutive accumulator columns [col0, col0 + 32) of output row
// `row` (crow = element offset of the row in C / residual), stored straight from registers.
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, const uint32_t (&r)[32], int row, int64_t crow, int col0) {
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

// split-K, every CTA: one thread's 32 accumulator columns of one row go to THIS split's private fp32 partial plane
// ws[split][M][N] with plain vector stores (round 1 and the first round-2 version added into one shared plane with
// red.global.add.f32: the L2 atomic units sustain only ~90 G elements/s, 35-60 us for a 2 M element output)
__device__ __forceinline__ void splitk_partial(const GemmParams& p, const uint32_t (&r)[32], int split, int row, int col0) {
  float* w = p.ws + ((int64_t)split * p.M + row) * p.N + col0;
  if ((p.N & 3) == 0) {
#pragma unroll
    for (int e = 0; e < 32; e += 4)
      if (col0 + e < p.N)
        __stcg(reinterpret_cast<float4*>(w + e), make_float4(__uint_as_float(r[e]), __uint_as_float(r[e + 1]),
                                                              __uint_as_float(r[e + 2]), __uint_as_float(r[e + 3])));
  } else {
#pragma unroll
    for (int e = 0; e < 32; ++e)
      if (col0 + e < p.N) __stcg(w + e, __uint_as_float(r[e]));
  }
}

// split-K, after the cluster barrier: split z of a tile applies out = act(alpha * sum_s ws[s] + bias + rowbias) + residual to ITS
// share (pieces [z, z + 1) * NP / splits) of the CTA's 128 x BN block.  te = 0..255 (the epilogue threads): one thread per
// (row, 8 columns) when N and ldc allow 16-byte accesses, else one per element.  The planes were written by other SMs: reads
// bypass L1.  Row-bias / residual loads are issued before the plane sums so that one round trip covers them all.
template <int BN>
__device__ __forceinline__ void splitk_finalize(const GemmParams& p, int m0, int n0, int z, int te, float* sstat) {
  const bool vec = (p.N % 8) == 0 && (p.ldc % 8) == 0 && (!p.rowbias || (p.rowbias_ld % 8) == 0);
  const int64_t plane = (int64_t)p.M * p.N;
  // GroupNorm statistics (fp16 output only; the host refuses other combinations): this split's rows fall into at most
  // two row groups (stats_rpg is a multiple of 128, or 64): shared-memory accumulators [2 groups][2 moments][BN], then one
  // red.add per column, moment and group for the whole share.
  const bool stats = p.colstats != nullptr && vec && !p.out_f32;
  if (stats) {
    for (int i = te; i < 4 * BN; i += EPI_THREADS) sstat[i] = 0.f;
    epi_bar();
  }
  if (vec) {
    constexpr int PPR = BN / 8;
    constexpr int NP = BM * PPR;
    const int share = (NP + p.splits - 1) / p.splits;
    const int i1 = min(NP, (z + 1) * share);
    // a thread's pieces are EPI_THREADS apart: when that is a multiple of the pieces per row its columns never change and
    // the bias is fetched once
    constexpr bool FIXED_COLS = (EPI_THREADS % PPR) == 0;
    float bcol[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) bcol[e] = 0.f;
    if (FIXED_COLS && p.bias) {
      const int i0 = z * share + te, col = n0 + (i0 - (i0 / PPR) * PPR) * 8;
      if (col < p.N) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col)), b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4));
        bcol[0] = b0.x, bcol[1] = b0.y, bcol[2] = b0.z, bcol[3] = b0.w, bcol[4] = b1.x, bcol[5] = b1.y, bcol[6] = b1.z, bcol[7] = b1.w;
      }
    }
    for (int i = z * share + te; i < i1; i += EPI_THREADS) {
      const int rl = i / PPR, row = m0 + rl, col = n0 + (i - rl * PPR) * 8;
      if (row >= p.M || col >= p.N) continue;
      const int64_t woff = (int64_t)row * p.N + col, o = out_row(p, row) * p.ldc + col;
      uint4 qb = make_uint4(0u, 0u, 0u, 0u), qr = make_uint4(0u, 0u, 0u, 0u);
      if (p.rowbias) qb = *reinterpret_cast<const uint4*>(p.rowbias + (int64_t)(row / p.rows_per_group) * p.rowbias_ld + col);
      if (p.residual) qr = *reinterpret_cast<const uint4*>(p.residual + o);
      if (!FIXED_COLS && p.bias) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col)), b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4));
        bcol[0] = b0.x, bcol[1] = b0.y, bcol[2] = b0.z, bcol[3] = b0.w, bcol[4] = b1.x, bcol[5] = b1.y, bcol[6] = b1.z, bcol[7] = b1.w;
      }
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 0.f;
      for (int s = 0; s < p.splits; s += 4) {
        float4 a[4], b[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const bool on = s + t < p.splits;
          const float4* w = reinterpret_cast<const float4*>(p.ws + (int64_t)(s + t) * plane + woff);
          a[t] = on ? __ldcg(w) : make_float4(0.f, 0.f, 0.f, 0.f);
          b[t] = on ? __ldcg(w + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          v[0] += a[t].x, v[1] += a[t].y, v[2] += a[t].z, v[3] += a[t].w;
          v[4] += b[t].x, v[5] += b[t].y, v[6] += b[t].z, v[7] += b[t].w;
        }
      }
      const __half* hb = reinterpret_cast<const __half*>(&qb);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float x = fmaf(v[e], p.alpha, p.rowbias ? __half2float(hb[e]) : 0.f) + bcol[e];
        v[e] = apply_act(x, p.act);
      }
      if (p.residual) {
        const __half* h = reinterpret_cast<const __half*>(&qr);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] += __half2float(h[e]);
      }
      store8(p, o, v);
      if (stats) {
        float* acc = sstat + (row / p.stats_rpg - m0 / p.stats_rpg) * 2 * BN + (col - n0);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float x = __half2float(__float2half_rn(v[e]));       // the value the consumer will read back
          atomicAdd(acc + e, x), atomicAdd(acc + BN + e, x * x);
        }
      }
    }
    if (stats) {
      epi_bar();
      const int g0 = m0 / p.stats_rpg;
      const int last_row = m0 + BM - 1 < p.M ? m0 + BM - 1 : p.M - 1;
      const int ng = last_row / p.stats_rpg - g0 + 1;                 // 1 or 2
      for (int i = te; i < ng * 2 * BN; i += EPI_THREADS) {
        const int gi = i / (2 * BN), rem = i - gi * 2 * BN, mom = rem / BN, c = rem - mom * BN;
        if (n0 + c < p.N && sstat[i] != 0.f) atomicAdd(p.colstats + ((int64_t)(g0 + gi) * 2 + mom) * p.N + n0 + c, sstat[i]);
      }
    }
    return;
  }
  const int share = (BM * BN + p.splits - 1) / p.splits;
  const int i1 = min(BM * BN, (z + 1) * share);
  for (int i = z * share + te; i < i1; i += EPI_THREADS) {
    const int rl = i / BN, row = m0 + rl, col = n0 + (i - rl * BN);
    if (row >= p.M || col >= p.N) continue;
    float x = 0.f;
    for (int s = 0; s < p.splits; ++s) x += __ldcg(p.ws + (int64_t)s * plane + (int64_t)row * p.N + col);
    x *= p.alpha;
    if (p.rowbias) x += __half2float(p.rowbias[(int64_t)(row / p.rows_per_group) * p.rowbias_ld + col]);
    if (p.bias) x += __ldg(p.bias + col);
    x = apply_act(x, p.act);
    const int64_t o = out_row(p, row) * p.ldc + col;
    if (p.residual) x += __half2float(p.residual[o]);
    if (p.out_f32) reinterpret_cast<float*>(p.C)[o] = x;
    else reinterpret_cast<__half*>(p.C)[o] = __float2half_rn(x);
  }
}

constexpr int EPI_RB_GROUPS = 8;                 // row-bias groups (images) one 128-row tile may span when staged in smem
__host__ __device__ constexpr int epi_smem_bytes(int bn) { return bn * 4 + EPI_RB_GROUPS * bn * 2; }
// the two column ranges of the eight epilogue warps split the tile at a multiple of 32 (a GEGLU chunk never straddles them)
__host__ __device__ constexpr int col_split(int bn) { return ((bn / 2 + 31) / 32) * 32; }

// Called by the eight epilogue warps while the main loop runs: the tile's bias and row-group-bias slices go to shared
// memory, so that the epilogue proper never waits on a first-touch global load.
template <int BN>
__device__ __forceinline__ void epilogue_preload(const GemmParams& p, int m0, int n0, float* sbias, __half* srb, int te) {
  for (int i = te; i < BN; i += EPI_THREADS) sbias[i] = (p.bias && n0 + i < p.N) ? __ldg(p.bias + n0 + i) : 0.f;
  if (p.rowbias) {
    const int g0 = m0 / p.rows_per_group;
    const int last = (m0 + BM - 1 < p.M ? m0 + BM - 1 : p.M - 1) / p.rows_per_group;
    const int ng = last - g0 + 1;
    if (ng >= 1 && ng <= EPI_RB_GROUPS)
      for (int i = te; i < ng * BN; i += EPI_THREADS) {
        const int gi = i / BN, c = i - gi * BN;
        srb[i] = (n0 + c < p.N) ? p.rowbias[(int64_t)(g0 + gi) * p.rowbias_ld + n0 + c] : __float2half(0.f);
      }
  }
  epi_bar();
}

template <int ACT>
__device__ __forceinline__ float act_fn(float x) {
  if (ACT == 1) return __fdividef(x, 1.f + __expf(-x));
  if (ACT == 2) return gelu_erf(x);
  if (ACT == 4) return __fdividef(x, 1.f + __expf(-1.702f * x));
  return x;
}

// Phase 1 of the staged epilogue for one thread (= one accumulator row) over tile columns [c_lo, c_hi): TMEM -> registers ->
// alpha / bias / row-group bias / activation (ACT 3: GEGLU gate) -> fp16 -> this row of the warp's shared-memory slab.
// rb: this row's row-group bias indexed by TILE column (nullptr: none); rb_smem: it is the shared-memory copy (no N tail).
template <int ACT>
__device__ __forceinline__ void stage_rows(const GemmParams& p, uint32_t tmem_row_base, uint8_t* mine, int n0, int c_lo, int c_hi,
                                           const float* sbias, const __half* rb, bool rb_smem) {
#pragma unroll 1
  for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
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
        h[e >> 1] = __floats2half2_rn(a0 * gelu_erf_fast(g0), a1 * gelu_erf_fast(g1));
      }
      uint4* d = reinterpret_cast<uint4*>(mine + (c0 - c_lo));   // (c0 - c_lo) / 2 output columns x 2 bytes
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
        *reinterpret_cast<uint4*>(mine + (c0 - c_lo + j) * 2) = *reinterpret_cast<uint4*>(h);
      }
    }
  }
}

// Output geometry of one epilogue warp in the staged modes: 32 rows x tile columns [c_lo, c_hi) become `outc` fp16 columns
// (GEGLU halves them) = `ppr` 16-byte pieces per row, `total` pieces per warp, lane l owns pieces l, l + 32, ...
struct WarpOut {
  int outc, stride, ppr, total, ocol0, nout;
};
template <int MODE>
__device__ __forceinline__ WarpOut warp_out(const GemmParams& p, int n0, int c_lo, int c_hi) {
  constexpr bool geglu = MODE == 1;
  WarpOut g;
  g.outc = geglu ? (c_hi - c_lo) >> 1 : (c_hi - c_lo);
  g.stride = g.outc * 2 + 16;                      // +16: 16-byte pieces of consecutive rows rotate banks
  g.ppr = g.outc >> 3;
  g.total = 32 * g.ppr;
  g.ocol0 = geglu ? (n0 + c_lo) >> 1 : n0 + c_lo;  // first column of C this warp writes
  g.nout = geglu ? p.N >> 1 : p.N;
  return g;
}

// The first RES_PREFETCH residual pieces of this lane, fetched while the main loop runs.
__device__ __forceinline__ void prefetch_residual(const GemmParams& p, const WarpOut& g, int lane, int row0,
                                                  uint4 (&resq)[RES_PREFETCH]) {
#pragma unroll
  for (int u = 0; u < RES_PREFETCH; ++u) {
    resq[u] = make_uint4(0u, 0u, 0u, 0u);
    const int pp = lane + 32 * u;
    if (p.residual && pp < g.total) {
      const int rl = pp / g.ppr, ci = pp - rl * g.ppr;
      const int grow = row0 + rl, col = g.ocol0 + ci * 8;
      if (grow < p.M && col < g.nout) resq[u] = *reinterpret_cast<const uint4*>(p.residual + out_row(p, grow) * p.ldc + col);
    }
  }
}

__device__ __forceinline__ uint4 add_h8(uint4 v, uint4 q) {
  __half2* a = reinterpret_cast<__half2*>(&v);
  const __half2* b = reinterpret_cast<const __half2*>(&q);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float2 fa = __half22float2(a[e]), fb = __half22float2(b[e]);
    a[e] = __floats2half2_rn(fa.x + fb.x, fa.y + fb.y);
  }
  return v;
}

// Staged epilogue of ONE warp: 32 accumulator rows (its TMEM lane quarter) x tile columns [c_lo, c_hi), fp16 output.
//   phase 1   thread = row, as the TMEM load delivers it: alpha / bias / row-group bias / activation / GEGLU gate, rounded to
//             fp16 (where autocast rounds the layer output) into this warp's slab of the (now idle) operand ring;
//   phase 2   the warp walks the slab in 16-byte pieces along rows, adds the residual (prefetched for the first
//             RES_PREFETCH pieces of a lane) and writes whole rows: every global access is a run of full 32-byte sectors.
// (r1 trace: with one 16-byte store per lane to 32 different rows the epilogue took 46 000 cycles per tile.)
template <int BN, int MODE>
__device__ __forceinline__ void epilogue_staged(const GemmParams& p, const WarpOut& g, uint32_t tmem_row_base, uint8_t* slab, int lane,
                                                int m0, int row0, int n0, int c_lo, int c_hi, const float* sbias, const __half* srb,
                                                const uint4 (&resq)[RES_PREFETCH], uint32_t release_bar = 0) {
  const int row = row0 + lane;
  // row-group bias of this thread's row: from the smem copy when the tile spans few groups, else straight from global
  const __half* rb = nullptr;
  bool rb_smem = false;
  if (MODE != 1 && p.rowbias) {
    const int rr = row < p.M ? row : p.M - 1;
    const int g0 = m0 / p.rows_per_group, last = (m0 + BM - 1 < p.M ? m0 + BM - 1 : p.M - 1) / p.rows_per_group;
    rb_smem = last - g0 + 1 <= EPI_RB_GROUPS;
    rb = rb_smem ? srb + (rr / p.rows_per_group - g0) * BN : p.rowbias + (int64_t)(rr / p.rows_per_group) * p.rowbias_ld + n0;
  }
  uint8_t* mine = slab + lane * g.stride;
  if (MODE == 0) {
    stage_rows<0>(p, tmem_row_base, mine, n0, c_lo, c_hi, sbias, rb, rb_smem);
  } else if (MODE == 1) {
    stage_rows<3>(p, tmem_row_base, mine, n0, c_lo, c_hi, sbias, rb, rb_smem);
  } else {   // the activation switch is hoisted out of the element loops: one branch per tile
    switch (p.act) {
      case 1: stage_rows<1>(p, tmem_row_base, mine, n0, c_lo, c_hi, sbias, rb, rb_smem); break;
      case 2: stage_rows<2>(p, tmem_row_base, mine, n0, c_lo, c_hi, sbias, rb, rb_smem); break;
      default: stage_rows<4>(p, tmem_row_base, mine, n0, c_lo, c_hi, sbias, rb, rb_smem); break;
    }
  }
  if (release_bar) {   // persistent kernel: this warp has read its part of the accumulator buffer -- hand it back to the MMA issuer
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(release_bar) : "memory");
  }
  __syncwarp();
  __half* C = reinterpret_cast<__half*>(p.C);
  // N is a multiple of 8 on this path (host check): no partial pieces
#pragma unroll
  for (int u = 0; u < RES_PREFETCH; ++u) {
    const int pp = lane + 32 * u;
    if (pp < g.total) {
      const int rl = pp / g.ppr, ci = pp - rl * g.ppr;
      const int grow = row0 + rl, col = g.ocol0 + ci * 8;
      if (grow < p.M && col < g.nout) {
        uint4 v = *reinterpret_cast<const uint4*>(slab + rl * g.stride + ci * 16);
        if (p.residual) {
          v = add_h8(v, resq[u]);
          if (p.colstats) *reinterpret_cast<uint4*>(slab + rl * g.stride + ci * 16) = v;   // the statistics see the final value
        }
        *reinterpret_cast<uint4*>(C + out_row(p, grow) * p.ldc + col) = v;
      }
    }
  }
  constexpr int UN = 4;
  for (int base = lane + 32 * RES_PREFETCH; base < g.total; base += 32 * UN) {
    uint4 v[UN], q[UN];
    int64_t o[UN];
    bool ok[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int pp = base + 32 * u;
      const int rl = pp / g.ppr, ci = pp - rl * g.ppr;
      const int grow = row0 + rl, col = g.ocol0 + ci * 8;
      ok[u] = pp < g.total && grow < p.M && col < g.nout;
      o[u] = out_row(p, grow) * p.ldc + col;
      if (ok[u]) {
        v[u] = *reinterpret_cast<const uint4*>(slab + rl * g.stride + ci * 16);
        if (p.residual) q[u] = *reinterpret_cast<const uint4*>(p.residual + o[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      if (!ok[u]) continue;
      if (p.residual) {
        v[u] = add_h8(v[u], q[u]);
        if (p.colstats) {
          const int pp = base + 32 * u;
          const int rl = pp / g.ppr, ci = pp - rl * g.ppr;
          *reinterpret_cast<uint4*>(slab + rl * g.stride + ci * 16) = v[u];
        }
      }
      *reinterpret_cast<uint4*>(C + o[u]) = v[u];
    }
  }
  // GroupNorm statistics of the tensor just written (the consumer's GroupNorm needs sum and sum of squares per image and
  // channel group): column sums over this warp's 32 rows -- all in one image, stats_rpg is a multiple of 32 -- read back
  // from the slab (lanes walk columns, conflict-free), one red.add per column and moment.
  if (MODE == 0 && p.colstats) {
    __syncwarp();
    const int nrows = p.M - row0 < 32 ? p.M - row0 : 32;
    if (nrows > 0) {
      float* ssum = p.colstats + (int64_t)(row0 / p.stats_rpg) * 2 * g.nout;
      for (int cp = lane; 2 * cp < g.outc; cp += 32) {
        const int col = g.ocol0 + 2 * cp;
        if (col >= g.nout) break;
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
        if (nrows == 32) {
#pragma unroll 8
          for (int r = 0; r < 32; ++r) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(slab + r * g.stride + cp * 4));
            s0 += f.x, s1 += f.y, q0 = fmaf(f.x, f.x, q0), q1 = fmaf(f.y, f.y, q1);
          }
        } else {
          for (int r = 0; r < nrows; ++r) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(slab + r * g.stride + cp * 4));
            s0 += f.x, s1 += f.y, q0 = fmaf(f.x, f.x, q0), q1 = fmaf(f.y, f.y, q1);
          }
        }
        atomicAdd(ssum + col, s0), atomicAdd(ssum + col + 1, s1);
        atomicAdd(ssum + g.nout + col, q0), atomicAdd(ssum + g.nout + col + 1, q1);
      }
    }
  }
}

__host__ __device__ constexpr int tmem_cols(int bn) { return bn <= 32 ? 32 : bn <= 64 ? 64 : bn <= 128 ? 128 : bn <= 256 ? 256 : 512; }
__host__ __device__ constexpr int epi_slab_bytes(int bn) { return EPI_WARPS * 32 * (col_split(bn) * 2 + 16); }
constexpr int smem_bytes(int bn, int stages, int ctas) {
  return stages * (BM * BK * 2 + (bn / ctas) * BK * 2) + (2 * stages + 1) * 8 + 8 + 16 + epi_smem_bytes(bn) + 1024;
}

// One stage of the TMA producer: this CTA's 128 rows of A and its BROWS rows of the B tile for k-block kb of the tile at
// (m0, n0).  Pair (CTAS = 2): the bytes of both CTAs are counted on the leader's full barrier.
template <int CTAS, int BROWS>
__device__ __forceinline__ void produce_stage(const GemmParams& p, const CUtensorMap* tmA, const CUtensorMap* tmB, uint8_t* a_dst,
                                              uint8_t* b_dst, uint64_t* full_bar, int kb, int m0, int n0, uint32_t rank, int bz) {
  constexpr int A_BYTES = BM * BK * 2, B_BYTES = BROWS * BK * 2;
  if (CTAS == 2) {
    if (rank == 0) mbar_expect_tx(full_bar, 2 * (A_BYTES + B_BYTES));   // the peer's bytes land on this barrier too
    const int nb = n0 + (int)rank * BROWS;
    if (p.conv) {
      const int tap = kb / p.cblocks, c0 = (kb - tap * p.cblocks) * BK;
      const int x0 = m0 % p.cW, y0 = (m0 / p.cW) % p.cH, b0 = m0 / (p.cW * p.cH);
      tma2_load_4d(a_dst, tmA, full_bar, c0, x0 + tap % p.ctx + p.cox, y0 + tap / p.ctx + p.coy, b0);
      tma2_load_2d(b_dst, tmB, full_bar, tap * p.cC + c0, nb);
    } else {
      tma2_load_2d(a_dst, tmA, full_bar, kb * BK, m0);
      tma2_load_2d(b_dst, tmB, full_bar, kb * BK, nb);
    }
  } else {
    mbar_expect_tx(full_bar, A_BYTES + B_BYTES);
    if (p.conv) {
      const int tap = kb / p.cblocks, c0 = (kb - tap * p.cblocks) * BK;
      const int x0 = m0 % p.cW, y0 = (m0 / p.cW) % p.cH, b0 = m0 / (p.cW * p.cH);
      tma_load_4d(a_dst, tmA, full_bar, c0, x0 + tap % p.ctx + p.cox, y0 + tap / p.ctx + p.coy, b0);
      tma_load_2d(b_dst, tmB, full_bar, tap * p.cC + c0, n0);
    } else if (p.batched) {
      tma_load_4d(a_dst, tmA, full_bar, kb * BK, m0, bz % p.nh, bz / p.nh);
      tma_load_4d(b_dst, tmB, full_bar, kb * BK, n0, bz % p.nh, bz / p.nh);
    } else {
      tma_load_2d(a_dst, tmA, full_bar, kb * BK, m0);
      tma_load_2d(b_dst, tmB, full_bar, kb * BK, n0);
    }
  }
}

// ------------------------------------------------------------------------------------------------ the kernel
// Two CTAs per SM: their prologues / epilogues overlap each other's main loops.  (Measured alternative, dropped: one CTA per
// SM with a ring twice as deep -- no gain on launches of <= 148 CTAs, 35 % slower on multi-wave launches: the main loops are
// not bound by bytes in flight; ncu shows 5.9 TB/s of L2 -> SM operand traffic on the 8192 x 320 x 2880 conv.)
template <int BN, int STAGES, int CTAS, int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by SWIZZLE_128B
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int BROWS = BN / CTAS;                           // rows of B staged by this CTA (pair: half of the tile)
  constexpr int A_BYTES = BM * BK * 2, B_BYTES = BROWS * BK * 2;
  static_assert(B_BYTES % 1024 == 0, "stage bases must stay 1024-byte aligned for SWIZZLE_128B");
  static_assert(BN % 32 == 0 && BN >= 64 && BN <= 256, "tile width: a multiple of 32 (GEGLU chunks, 32-column TMEM loads)");
  static_assert(epi_slab_bytes(BN) <= STAGES * (A_BYTES + B_BYTES), "the output transpose must fit in the operand ring");
  constexpr int TCOLS = tmem_cols(BN);
  constexpr int CSPLIT = col_split(BN);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);   // [0] TMEM base address, [1] split-K "this CTA is the last"
  float* sbias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 2) + 15) & ~(uintptr_t)15);
  __half* srb = reinterpret_cast<__half*>(sbias + BN);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) stamp(p, 0), stamp_ns(p, 16);
  // cluster = (CTAS, 1, splits): ranks 2j and 2j + 1 are the pair of split j.  rank 0 = leader of its pair (issues the MMAs,
  // owns the full barriers)
  const uint32_t crank = cluster_ctarank();
  const uint32_t rank = CTAS == 2 ? (crank & 1u) : 0u;
  const uint16_t pair_mask = (uint16_t)(3u << (crank & ~1u));
  const int m0 = CTAS == 2 ? (blockIdx.x >> 1) * (2 * BM) + (int)rank * BM : blockIdx.x * BM;
  const int n0 = blockIdx.y * BN, bz = blockIdx.z;
  const int nk = p.conv ? p.ctaps * p.cblocks : (p.K + BK - 1) / BK;
  // split-K: blockIdx.z owns k-blocks [kb0, kb1) and adds its partial tile into the fp32 workspace
  int kb0 = 0, kb1 = nk;
  if (MODE == 3 && p.splits > 1) {
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
  if (warp == 1) {  // TMEM: TCOLS fp32 columns x 128 lanes (pair: the same warp of BOTH CTAs allocates the pair's columns)
    if (CTAS == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TCOLS));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TCOLS));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (CTAS == 2) cluster_sync_all();   // barrier inits of the leader must be visible before the peer's TMA can complete on them
  else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // everything above touched no global memory: it overlaps the tail of the previous kernel
  pdl_trigger();
  if (threadIdx.x == 0) stamp(p, 1);

  if (warp == 0) {
    if (lane == 0) {  // ---------------- TMA producer: own 128 rows of A, own BROWS rows of the B tile
      for (int kb = kb0; kb < kb1; ++kb) {
        const int s = (kb - kb0) % STAGES;
        const uint32_t ph = ((kb - kb0) / STAGES) & 1;
        mbar_wait(empty + s, ph ^ 1, p, WAIT_EMPTY, s);
        produce_stage<CTAS, BROWS>(p, &tmA, &tmB, sA + s * A_BYTES, sB + s * B_BYTES, full + s, kb, m0, n0, rank, bz);
        if (kb == kb0) stamp(p, 2);
      }
      stamp(p, 3);
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {  // ---------------- MMA issuer (pair: the leader only, M = 256 across the pair)
      constexpr uint32_t idesc = umma_idesc_f16(CTAS * BM, BN);
      for (int kb = kb0; kb < kb1; ++kb) {
        const int s = (kb - kb0) % STAGES;
        const uint32_t ph = ((kb - kb0) / STAGES) & 1;
        mbar_wait(full + s, ph, p, WAIT_FULL, s);
        if (kb == kb0) stamp(p, 4);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a0 = smem_u32(sA + s * A_BYTES), b0 = smem_u32(sB + s * B_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)   // advancing 16 fp16 along K inside the 128-byte swizzle atom = +32 bytes on the start address
          umma_f16<CTAS>(tmem_base, umma_desc_sw128(a0 + k * 32), umma_desc_sw128(b0 + k * 32), idesc, ((kb - kb0) | k) != 0);
        umma_commit<CTAS>(empty + s, pair_mask);   // frees this stage (pair: in both CTAs) once these MMAs have read it
      }
      umma_commit<CTAS>(tmem_full, pair_mask);     // accumulator complete: the epilogue warps (pair: of both CTAs) may start
      stamp(p, 5);
    }
  } else {  // ------------------------ epilogue warps 2..9 on this CTA's 128 accumulator rows
    const int e = warp - 2, quarter = warp & 3, te = threadIdx.x - 64;
    const int c_lo = e < 4 ? 0 : CSPLIT, c_hi = e < 4 ? CSPLIT : BN;   // this warp's tile columns
    const int row0 = m0 + quarter * 32;
    const uint32_t tmem_row_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    if (MODE != 3) {
      epilogue_preload<BN>(p, m0, n0, sbias, srb, te);
      const WarpOut g = warp_out<MODE>(p, n0, c_lo, c_hi);
      uint4 resq[RES_PREFETCH];
      prefetch_residual(p, g, lane, row0, resq);
      mbar_wait(tmem_full, 0, p, WAIT_ACC, 0);
      if (threadIdx.x == 64) stamp(p, 6);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // the operand ring is idle once the accumulator is complete: it becomes the output transpose buffer
      epilogue_staged<BN, MODE>(p, g, tmem_row_base, smem + e * 32 * (CSPLIT * 2 + 16), lane, m0, row0, n0, c_lo, c_hi, sbias, srb, resq);
    } else {
      mbar_wait(tmem_full, 0, p, WAIT_ACC, 0);
      if (threadIdx.x == 64) stamp(p, 6);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int row = row0 + lane;
      if (p.splits > 1) {
        if (te == 0) stamp_ns(p, 17);
#pragma unroll 1
        for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_row_base + c0, r);
          if (row < p.M) splitk_partial(p, r, bz, row, n0 + c0);
        }
        if (te == 0) stamp_ns(p, 18);
      } else {
        const int64_t crow = (p.batched ? (int64_t)(bz % p.nh) * p.stride_c_h + (int64_t)(bz / p.nh) * p.stride_c_b : 0) +
                             out_row(p, row < p.M ? row : 0) * p.ldc;
#pragma unroll 1
        for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_row_base + c0, r);
          if (row < p.M) epilogue_chunk(p, r, row, crow, n0 + c0);
        }
      }
    }
    if (threadIdx.x == 64) stamp(p, 7);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  // pair: neither CTA may free TMEM / exit while the pair's MMAs or the peer's TMEM reads are in flight.
  // split-K: the `splits` CTAs (pairs) of a tile form ONE cluster (co-scheduled by the hardware), so this barrier -- release /
  // acquire at cluster scope -- also publishes every split's partial plane to its siblings.
  const bool split = MODE == 3 && p.splits > 1;
  if (CTAS == 2 || split) cluster_sync_all();
  else __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (CTAS == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TCOLS));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TCOLS));
  }
  if (MODE == 3) {
    if (split && warp >= 2) {   // every split finalizes its share of the 128 x BN block: sum of the planes + epilogue
      if (threadIdx.x == 64) stamp_ns(p, 21, true);
      splitk_finalize<BN>(p, m0, n0, bz, threadIdx.x - 64, reinterpret_cast<float*>(smem));
      if (threadIdx.x == 64) stamp_ns(p, 22, true);
    }
  }
  if (threadIdx.x == 0) stamp(p, 8);
}


// ------------------------------------------------------------------------------------------------ the persistent kernel
// Problems with MANY tiles (the batched sampler calls: M = 16 384 ... 65 536 rows): one CTA pair per SM pair walks tiles
// pair, pair + n_pairs, ... with TWO accumulator buffers in TMEM, so that
//   * the epilogue of tile i (TMEM -> registers -> bias / activation / GEGLU -> shared-memory transpose -> global) runs
//     under the main loop of tile i + 1 -- on short-K problems (K = 320: five k-blocks) the epilogue is as long as the main
//     loop, and a CTA per tile serialises them (two resident CTAs per SM hide only part of it: 190 us against cuBLAS's 103
//     on the 65536 x 2560 x 320 GEGLU projection);
//   * the producer runs ahead across tile borders (the ring never drains), and barrier set-up, TMEM allocation, tensor-map
//     fetch and tear-down are paid once per SM instead of once per tile.
// Accumulator hand-over: acc_full[b] (tcgen05.commit, multicast to the pair) MMA -> epilogue warps of both CTAs;
// acc_free[b] on the LEADER, 16 arrivals (eight epilogue warps x two CTAs, remote arrive from the peer) epilogue -> MMA.
// The output transpose has its own shared memory (the ring is never idle), bias / row-bias slices are double-buffered.
// Pair tiles only (CTAS = 2), staged fp16 epilogues only (MODE 0 / 1 / 2): everything else takes gemm_tc_kernel.
constexpr int persist_smem_bytes(int bn, int stages) {
  return stages * (BM * BK * 2 + (bn / 2) * BK * 2) + epi_slab_bytes(bn) + (2 * stages + 4) * 8 + 16 + 2 * epi_smem_bytes(bn) + 1024;
}

template <int BN, int STAGES, int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                          const __grid_constant__ GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int BROWS = BN / 2;
  constexpr int A_BYTES = BM * BK * 2, B_BYTES = BROWS * BK * 2;
  static_assert(B_BYTES % 1024 == 0, "stage bases must stay 1024-byte aligned for SWIZZLE_128B");
  static_assert(MODE != 3, "staged epilogues only");
  constexpr int TCOLS = tmem_cols(BN);                       // one accumulator buffer; two are allocated
  static_assert(2 * TCOLS <= 512, "two accumulator buffers must fit the 512 TMEM columns");
  constexpr int CSPLIT = col_split(BN);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint8_t* slabs = sB + STAGES * B_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(slabs + epi_slab_bytes(BN));
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint64_t* acc_free = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 2);
  float* sbias0 = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 2) + 15) & ~(uintptr_t)15);
  constexpr int EPI_FLOATS = epi_smem_bytes(BN) / 4;         // one bias + row-bias buffer, in floats

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const uint32_t rank = crank & 1u;
  const uint16_t pair_mask = (uint16_t)(3u << (crank & ~1u));
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int tiles_n = (p.N + BN - 1) / BN;
  const int ntiles = ((p.M + 2 * BM - 1) / (2 * BM)) * tiles_n;
  const int nk = p.conv ? p.ctaps * p.cblocks : (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < STAGES; ++s) mbar_init(full + s, 1), mbar_init(empty + s, 1);
    for (int b = 0; b < 2; ++b) mbar_init(acc_full + b, 1), mbar_init(acc_free + b, 2 * EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * TCOLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    if (lane == 0) {  // ---------------- TMA producer: runs ahead across tile borders, bounded only by the ring
      uint32_t cnt = 0;
      for (int t = pair; t < ntiles; t += npairs) {
        const int tm = t / tiles_n, n0 = (t - tm * tiles_n) * BN, m0 = tm * (2 * BM) + (int)rank * BM;
        for (int kb = 0; kb < nk; ++kb, ++cnt) {
          const int s = cnt % STAGES;
          mbar_wait(empty + s, ((cnt / STAGES) & 1) ^ 1, p, WAIT_EMPTY, s);
          produce_stage<2, BROWS>(p, &tmA, &tmB, sA + s * A_BYTES, sB + s * B_BYTES, full + s, kb, m0, n0, rank, 0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {  // ---------------- MMA issuer (leader CTA): tile i into accumulator buffer i & 1
      constexpr uint32_t idesc = umma_idesc_f16(2 * BM, BN);
      uint32_t cnt = 0, it = 0;
      for (int t = pair; t < ntiles; t += npairs, ++it) {
        const uint32_t buf = it & 1;
        mbar_wait(acc_free + buf, ((it >> 1) & 1) ^ 1, p, WAIT_ACC_FREE, (int)buf);   // both CTAs' epilogue warps have drained tile it - 2
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t acc = tmem_base + buf * TCOLS;
        for (int kb = 0; kb < nk; ++kb, ++cnt) {
          const int s = cnt % STAGES;
          mbar_wait(full + s, (cnt / STAGES) & 1, p, WAIT_FULL, s);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a0 = smem_u32(sA + s * A_BYTES), b0 = smem_u32(sB + s * B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_f16<2>(acc, umma_desc_sw128(a0 + k * 32), umma_desc_sw128(b0 + k * 32), idesc, (kb | k) != 0);
          umma_commit<2>(empty + s, pair_mask);
        }
        umma_commit<2>(acc_full + buf, pair_mask);
      }
    }
  } else {  // ------------------------ epilogue warps 2..9: this CTA's 128 accumulator rows of every tile
    const int e = warp - 2, quarter = warp & 3, te = threadIdx.x - 64;
    const int c_lo = e < 4 ? 0 : CSPLIT, c_hi = e < 4 ? CSPLIT : BN;
    uint8_t* slab = slabs + e * 32 * (CSPLIT * 2 + 16);
    // the leader's acc_free barriers, as shared::cluster addresses (the peer arrives remotely)
    uint32_t free_bar0;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(free_bar0) : "r"(smem_u32(acc_free)), "r"(crank & ~1u));
    uint32_t it = 0;
    for (int t = pair; t < ntiles; t += npairs, ++it) {
      const uint32_t buf = it & 1;
      const int tm = t / tiles_n, n0 = (t - tm * tiles_n) * BN, m0 = tm * (2 * BM) + (int)rank * BM;
      const int row0 = m0 + quarter * 32;
      float* sbias = sbias0 + buf * EPI_FLOATS;
      __half* srb = reinterpret_cast<__half*>(sbias + BN);
      epilogue_preload<BN>(p, m0, n0, sbias, srb, te);
      const WarpOut g = warp_out<MODE>(p, n0, c_lo, c_hi);
      uint4 resq[RES_PREFETCH];
      prefetch_residual(p, g, lane, row0, resq);
      mbar_wait(acc_full + buf, (it >> 1) & 1, p, WAIT_ACC, (int)buf);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tmem_row_base = tmem_base + buf * TCOLS + ((uint32_t)(quarter * 32) << 16);
      epilogue_staged<BN, MODE>(p, g, tmem_row_base, slab, lane, m0, row0, n0, c_lo, c_hi, sbias, srb, resq, free_bar0 + buf * 8);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();   // neither CTA may free TMEM / exit while the pair's MMAs or the peer's TMEM reads are in flight
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * TCOLS));
  }
}

// ------------------------------------------------------------------------------------------------ host side
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

// Host-mapped trap record (one per process): allocated on the first launch that is not inside a stream capture.
TrapRecord* g_diag = nullptr;
bool g_diag_tried = false;
TrapRecord* diag_buffer(cudaStream_t st) {
  if (!g_diag_tried) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
      cudaGetLastError();
      return g_diag;
    }
    g_diag_tried = true;
    void* h = nullptr;
    if (cudaHostAlloc(&h, sizeof(TrapRecord), cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess) {
      memset(h, 0, sizeof(TrapRecord));
      g_diag = reinterpret_cast<TrapRecord*>(h);
    } else {
      cudaGetLastError();
    }
  }
  return g_diag;
}

struct Config {
  int ctas, bn, splits;
  int persist;   // 1: gemm_tc_persistent_kernel (pair tiles, staged epilogue, many tiles)
};

// tuning / sweep hook: O2345_GEMM_FORCE="ctas,bn,splits" (0 = keep the heuristic's choice for that field); also settable
// through o2345_debug_gemm_force (tools/gemm_sweep.py)
int g_force[3] = {-1, 0, 0};
int g_persist = 0;          // 0: heuristic, 1: persistent kernel wherever it is available, 2: never (o2345_debug_gemm_persist)
int g_persist_min_tiles = 0;   // heuristic threshold override (0: default)
void read_force_env() {
  if (g_force[0] >= 0) return;
  g_force[0] = g_force[1] = g_force[2] = 0;
  const char* e = getenv("O2345_GEMM_FORCE");
  if (e) sscanf(e, "%d,%d,%d", &g_force[0], &g_force[1], &g_force[2]);
}

// Tile shape and k-splits for a problem of M x N with nk k-blocks of 64: the candidate with the lowest predicted time under a
// small cost model fitted to tools/gemm_sweep.py (63 UNet shapes x ~40 forced configurations each on a B200, cold L2;
// profiles/r2_gemm_sweep.txt: the model's picks total 3.23 ms against 3.18 ms for the per-shape optimum):
//   * a launch costs ~5 us of fixed latency (launch, prologue, first TMA round trip, tear-down) + ~3 us of epilogue per
//     160 columns of tile and wave of 296 CTAs;
//   * the main loop is bound by operand delivery, not by the tensor pipe: an SM ingests ~40 B/clk (60 when two CTAs share it
//     and cover each other's bubbles), the whole L2 -> SM fabric ~6 300 B/clk -- so few fat tiles starve (few SMs pull),
//     many thin tiles re-read A (fabric), and long-K problems with few tiles want split-K;
//   * split-K adds ~4 us + ~1 us per split and 128 tile columns (partial planes through L2, cluster barrier, finalize).
float g_model[7] = {40.f, 0.5f, 6300.f, 5.f, 3.f, 4.f, 1.f};   // bw_sm, alpha, cap, fixed, epi, so0, so1 (tools: o2345_debug_gemm_model)
float predict_us(int M, int N, int nk, int ctas, int bn, int splits) {
  const float bw_sm = g_model[0], alpha = g_model[1], cap = g_model[2], fixed = g_model[3], epi = g_model[4], so0 = g_model[5],
              so1 = g_model[6], cyc_per_us = 1900.f;
  const int mblocks = ctas == 2 ? 2 * cdiv(M, 2 * BM) : cdiv(M, BM);
  const int tiles = mblocks * cdiv(N, bn);
  const int n = tiles * splits;
  const int kb = cdiv(nk, splits);
  const float bytes_cta = (float)kb * (float)(BM * BK * 2 + (bn / ctas) * BK * 2);
  const int sms = sm_count();
  const int rounds = cdiv(n, sms);
  float f = (float)(n - sms) / (float)sms;
  f = f < 0.f ? 0.f : (f > 1.f ? 1.f : f);
  const float t_sm = (float)rounds * bytes_cta / (bw_sm * (1.f + alpha * f));
  const float t_fabric = (float)n * bytes_cta / cap;
  const float t_mma = (float)rounds * (float)kb * 4.f * ((float)bn * 0.5f);
  float t = t_sm > t_fabric ? t_sm : t_fabric;
  if (t_mma > t) t = t_mma;
  float us = fixed + t / cyc_per_us + epi * (float)bn / 160.f * (float)cdiv(n, 2 * sms);
  if (splits > 1) us += so0 + so1 * (float)splits * (float)bn / 128.f;
  return us;
}

Config pick_config(const GemmParams& p, int nk, bool can_split, int64_t ws_floats) {
  read_force_env();
  Config c;
  const int M = p.M, N = p.N;
  c.persist = 0;
  if (p.batched) {
    c.ctas = 1, c.bn = N <= 64 ? 64 : 128, c.splits = 1;
    return c;
  }
  static const int kBn[4] = {64, 128, 160, 256};
  static const int kSplits[8] = {1, 2, 3, 4, 6, 8, 12, 16};
  const bool split_ok = can_split && p.act != 3 && p.ws && 2 * (int64_t)M * N <= ws_floats;
  const int max_planes = split_ok ? (int)(ws_floats / ((int64_t)M * N)) : 1;
  float best = 1e30f;
  c.ctas = 2, c.bn = 128, c.splits = 1;
  for (int ctas = 1; ctas <= 2; ++ctas) {
    if (ctas == 1 && M > 2 * BM) continue;                      // single-CTA tiles only pay for short problems
    if ((g_force[0] == 1 || g_force[0] == 2) && ctas != g_force[0]) continue;
    for (int bi = 0; bi < 4; ++bi) {
      const int bn = kBn[bi];
      if (ctas == 1 && bn > 128) continue;
      if (bn > 64 && N <= 64) continue;
      if (g_force[1] > 0 && bn != g_force[1]) continue;
      for (int si = 0; si < 8; ++si) {
        const int sp = kSplits[si];
        if (g_force[2] > 0 && sp != (g_force[2] > nk ? nk : g_force[2])) continue;
        if (sp > 1 && (!split_ok || sp > max_planes || sp * ctas > MAX_CLUSTER || nk / sp < 3)) continue;
        const float us = predict_us(M, N, nk, ctas, bn, sp);
        if (us < best) best = us, c.ctas = ctas, c.bn = bn, c.splits = sp;
      }
    }
  }
  if (best > 1e29f) {   // a forced configuration that is not available: fall back to the nearest valid one
    c.ctas = (g_force[0] == 1) ? 1 : 2;
    c.bn = (g_force[1] == 64 || g_force[1] == 128 || (c.ctas == 2 && (g_force[1] == 160 || g_force[1] == 256))) ? g_force[1] : 128;
    c.splits = 1;
  }
  // Many tiles and a short K (the batched sampler calls: M = 16 384 ... 65 536): the epilogue of a tile is as long as its main
  // loop, and the persistent kernel hides it under the next tile's.  Measured per shape (tools/gemm_persist_ab.py, B200):
  // 65536 x 2560 x 320 GEGLU 184 -> 157 us, 65536 x 960 x 320 75 -> 59, 65536 x 320 x 1280 85 -> 63, 16384 x 1920 x 640 51 -> 38;
  // with long K it LOSES (65536 x 320 x 2880 conv 111 -> 141 us): one CTA per SM keeps 128-156 KB of operands in flight
  // against 192-208 KB for two resident per-tile CTAs, and the main loop is bound by bytes in flight.
  if (g_persist == 0 && g_force[0] == 0 && g_force[1] == 0 && g_force[2] == 0 && !p.out_f32 && nk <= 20) {
    // tile width: 256 where it divides the work well, 160 for the UNet's N = 320 / 640 / 960, 128 for N = 128 (the VAE's top level)
    const int bn_p = (N >= 512 || N == 256) ? 256 : ((N % 160) == 0 || N > 128) ? 160 : 128;
    const int tiles = cdiv(M, 2 * BM) * cdiv(N, bn_p);
    const int min_tiles = g_persist_min_tiles > 0 ? g_persist_min_tiles : sm_count() / 2;
    if (N >= 128 && tiles >= min_tiles && (N >= 512 || nk >= 8)) c.ctas = 2, c.bn = bn_p, c.splits = 1, c.persist = 1;
  }
  return c;
}

// which compile-time epilogue: staged fp16 output needs whole 16-byte pieces and one pass (no split-K, no head batches)
int pick_mode(const GemmParams& p, const Config& c) {
  const int nout = p.act == 3 ? p.N / 2 : p.N;
  const bool staged = !p.out_f32 && c.splits <= 1 && !p.batched && (nout % 8) == 0 && (p.ldc % 8) == 0 && ((uintptr_t)p.C % 16) == 0 &&
                      (!p.residual || ((uintptr_t)p.residual % 16) == 0) && !(c.ctas == 1 && p.act == 3);
  if (!staged) return 3;
  return p.act == 0 ? 0 : (p.act == 3 ? 1 : 2);
}

// The persistent kernel (mode 0 / 1 / 2 epilogues, pair tiles): chosen by pick_config for many-tile short-K problems, or forced
// by the tuning hook wherever it is available.
bool pick_persist(const Config& c, int mode) {
  if (g_persist == 2 || mode == 3 || c.ctas != 2 || c.splits > 1 || c.bn < 128) return false;
  return g_persist == 1 || c.persist == 1;
}

long long* g_trace = nullptr;

template <int BN, int STAGES, int CTAS, int MODE>
int launch(const CUtensorMap& a, const CUtensorMap& b, GemmParams p, int batch, cudaStream_t st) {
  constexpr int SMEM = smem_bytes(BN, STAGES, CTAS);
  static_assert(2 * SMEM <= 227 * 1024 + 2048, "two CTAs per SM");
  static PerDeviceOnce attr;
  if (attr.need()) {
    O2345_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, CTAS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    if (MODE == 3)   // split-K clusters of up to 16 CTAs
      O2345_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, CTAS, MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  }
  p.bn = BN, p.ctas = CTAS, p.mode = MODE;
  p.diag = diag_buffer(st);
  const int mblocks = CTAS == 2 ? 2 * cdiv(p.M, 2 * BM) : cdiv(p.M, BM);
  const int splits = (MODE == 3 && batch == 0 && p.splits > 1) ? p.splits : 1;
  p.splits = splits;
  dim3 grid(mblocks, cdiv(p.N, BN), batch > 0 ? batch : splits);
  O2345_CUDA(launch_pdl_cluster(gemm_tc_kernel<BN, STAGES, CTAS, MODE>, grid, dim3(GEMM_THREADS), (size_t)SMEM, st, CTAS, splits, a, b, p));
  O2345_LAUNCH_CHECK();
  return O2345_OK;
}

template <int BN, int STAGES, int MODE>
int launch_persistent(const CUtensorMap& a, const CUtensorMap& b, GemmParams p, cudaStream_t st) {
  constexpr int SMEM = persist_smem_bytes(BN, STAGES);
  static_assert(SMEM <= 227 * 1024, "one CTA per SM");
  static PerDeviceOnce attr;
  if (attr.need())
    O2345_CUDA(cudaFuncSetAttribute(gemm_tc_persistent_kernel<BN, STAGES, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  p.bn = BN, p.ctas = 2, p.mode = MODE + 10;
  p.diag = diag_buffer(st);
  p.splits = 1;
  const int tiles = cdiv(p.M, 2 * BM) * cdiv(p.N, BN), pairs = sm_count() / 2;
  dim3 grid(2 * (tiles < pairs ? tiles : pairs), 1, 1);
  O2345_CUDA(launch_pdl_cluster(gemm_tc_persistent_kernel<BN, STAGES, MODE>, grid, dim3(GEMM_THREADS), (size_t)SMEM, st, 2, 1, a, b, p));
  O2345_LAUNCH_CHECK();
  return O2345_OK;
}

template <int BN, int STAGES>
int launch_persistent_mode(int mode, const CUtensorMap& a, const CUtensorMap& b, const GemmParams& p, cudaStream_t st) {
  switch (mode) {
    case 0: return launch_persistent<BN, STAGES, 0>(a, b, p, st);
    case 1: return launch_persistent<BN, STAGES, 1>(a, b, p, st);
    default: return launch_persistent<BN, STAGES, 2>(a, b, p, st);
  }
}

template <int BN, int STAGES, int CTAS>
int launch_mode(int mode, const CUtensorMap& a, const CUtensorMap& b, const GemmParams& p, int batch, cudaStream_t st) {
  switch (mode) {
    case 0: return launch<BN, STAGES, CTAS, 0>(a, b, p, batch, st);
    case 1:
      if constexpr (CTAS == 2) return launch<BN, STAGES, 2, 1>(a, b, p, batch, st);
      else return launch<BN, STAGES, CTAS, 3>(a, b, p, batch, st);
    case 2: return launch<BN, STAGES, CTAS, 2>(a, b, p, batch, st);
    default: return launch<BN, STAGES, CTAS, 3>(a, b, p, batch, st);
  }
}

int dispatch(const Config& c, int mode, const CUtensorMap& a, const CUtensorMap& b, const GemmParams& p, int batch, cudaStream_t st) {
  if (p.colstats && mode == 3 && c.splits <= 1) {
    set_error("o2345_gemm_f16: column statistics need the staged fp16 epilogue (16-byte aligned C / residual) or split-K");
    return O2345_EUNSUPPORTED;
  }
  if (pick_persist(c, mode)) {
    if (c.bn == 128) return launch_persistent_mode<128, 6>(mode, a, b, p, st);
    if (c.bn == 160) return launch_persistent_mode<160, 6>(mode, a, b, p, st);
    return launch_persistent_mode<256, 4>(mode, a, b, p, st);
  }
  if (c.ctas == 2) {
    if (c.bn == 64) return launch_mode<64, 5, 2>(mode, a, b, p, batch, st);
    if (c.bn == 128) return launch_mode<128, 4, 2>(mode, a, b, p, batch, st);
    if (c.bn == 160) return launch_mode<160, 4, 2>(mode, a, b, p, batch, st);
    return launch_mode<256, 3, 2>(mode, a, b, p, batch, st);
  }
  if (c.bn == 64) return launch_mode<64, 4, 1>(mode, a, b, p, batch, st);
  return launch_mode<128, 3, 1>(mode, a, b, p, batch, st);
}

int fill_epilogue(GemmParams& p, const o2345_epilogue* ep, int M, int N, int64_t ldc) {
  p.bias = nullptr, p.rowbias = nullptr, p.rowbias_ld = 0, p.rows_per_group = 1, p.residual = nullptr;
  p.out_f32 = 0, p.act = 0, p.alpha = 1.f, p.trace = g_trace;
  p.colstats = nullptr, p.stats_rpg = 1, p.stats_groups = 0;
  p.diag = nullptr, p.bn = p.ctas = p.mode = 0;
  if (!ep) return O2345_OK;
  O2345_CHECK_ARG(ep->act >= 0 && ep->act <= 4, "unknown activation");
  O2345_CHECK_ARG(!ep->rowbias || (ep->rows_per_group > 0 && (ep->rowbias_ld % 8) == 0 && ((uintptr_t)ep->rowbias % 16) == 0),
                  "row bias: rows_per_group > 0, 16-byte aligned, row stride a multiple of 8");
  O2345_CHECK_ARG(ep->act != 3 || ((N % 32) == 0 && (ldc % 8) == 0 && !ep->residual && !ep->rowbias),
                  "GEGLU epilogue: N must be a multiple of 32, ldc of 8, no residual / row bias");
  p.bias = ep->bias, p.rowbias = reinterpret_cast<const __half*>(ep->rowbias), p.rowbias_ld = ep->rowbias_ld;
  p.rows_per_group = ep->rowbias ? ep->rows_per_group : 1;
  p.residual = reinterpret_cast<const __half*>(ep->residual), p.out_f32 = ep->out_f32, p.act = ep->act, p.alpha = ep->alpha;
  if (ep->colstats) {
    O2345_CHECK_ARG(ep->act == 0 && !ep->out_f32 && (N % 8) == 0 && (ldc % 8) == 0,
                    "column statistics: fp16 output without activation, N and ldc multiples of 8");
    O2345_CHECK_ARG(ep->stats_rows_per_group > 0 && ((ep->stats_rows_per_group % 128) == 0 || ep->stats_rows_per_group == 64),
                    "column statistics: rows per group must be 64 or a multiple of 128");
    p.colstats = ep->colstats, p.stats_rpg = ep->stats_rows_per_group, p.stats_groups = cdiv(M, ep->stats_rows_per_group);
  }
  return O2345_OK;
}

void set_workspace(GemmParams& p, float* ws, int64_t ws_floats) { p.ws = ws_floats > 0 ? ws : nullptr; }

}  // namespace
}  // namespace o2345

using namespace o2345;

extern "C" void o2345_debug_gemm_trace(long long* device_buf16) { g_trace = device_buf16; }

extern "C" void o2345_debug_gemm_force(int ctas, int bn, int splits) {
  read_force_env();
  g_force[0] = ctas, g_force[1] = bn, g_force[2] = splits;
}


extern "C" void o2345_debug_gemm_persist(int mode, int min_tiles) { g_persist = mode, g_persist_min_tiles = min_tiles; }

extern "C" void o2345_debug_gemm_model(const float* seven) {
  for (int i = 0; i < 7; ++i) g_model[i] = seven[i];
}

extern "C" int o2345_last_trap(char* buf, size_t n) {
  if (!buf || n == 0) return O2345_EINVAL;
  buf[0] = 0;
  const TrapRecord* d = g_diag;
  if (!d || d->magic != TRAP_MAGIC) return 0;
  static const char* names[] = {"?", "empty (producer waiting for the MMA to free a stage)", "full (MMA issuer waiting for TMA bytes)",
                                "accumulator (epilogue waiting for the last MMA)",
                                "accumulator buffer (MMA issuer waiting for the epilogue warps to drain it)"};
  snprintf(buf, n,
           "gemm_tc_kernel<BN=%d, CTAS=%d, MODE=%d> M=%d N=%d K=%d conv=%d splits=%d: CTA (%d,%d,%d) rank %d gave up after 4 s "
           "on barrier '%s' stage %d",
           d->bn, d->ctas, d->mode, d->M, d->N, d->K, d->conv, d->splits, d->bx, d->by, d->bz, d->rank,
           names[d->tag >= 1 && d->tag <= 4 ? d->tag : 0], d->stage);
  return 1;
}

namespace o2345 {
namespace {
// One implicit convolution launch over the channel-last activation x [B, H, W, C]: `taps` taps in rows of `tx`, tap j shifted by
// (j % tx + ox, j / tx + oy); weight [N, taps * C] in (tap, channel) order; up / upa / upb: see GemmParams.
int conv_launch(const void* x, int B, int H, int W, int C, const void* weight, int N, void* out, int64_t ldc, const o2345_epilogue* ep,
                float* splitk_ws, int64_t ws_floats, int taps, int tx, int ox, int oy, int up, int upa, int upb, cudaStream_t st) {
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
  p.M = B * H * W, p.N = N, p.K = taps * C, p.ldc = ldc, p.nh = 1, p.stride_c_h = 0, p.stride_c_b = 0, p.C = out;
  int rc = fill_epilogue(p, ep, p.M, N, ldc);
  if (rc) return rc;
  p.batched = 0, p.conv = 1, p.cC = C, p.cH = H, p.cW = W, p.cblocks = (C + BK - 1) / BK;
  p.ctaps = taps, p.ctx = tx, p.cox = ox, p.coy = oy, p.up = up, p.upa = upa, p.upb = upb;
  set_workspace(p, splitk_ws, ws_floats);
  const Config c = pick_config(p, taps * p.cblocks, true, ws_floats);
  p.splits = c.splits;
  rc = make_map(&mb, weight, N, taps * (int64_t)C, taps * (int64_t)C, 0, 0, 0, 0, c.bn / c.ctas);
  if (rc) return rc;
  return dispatch(c, pick_mode(p, c), ma, mb, p, 0, st);
}

// an output tile is 128 consecutive pixels fetched as ONE box (tw, th, tb): either whole multiples of 128 along a row,
// or whole rows that tile the image exactly (H a multiple of 128 / W), or whole images (H * W divides 128).  Any other
// shape would wrap a tile across the image border (silently wrong rows) or give a box of fewer than 128 rows (the
// stage's byte count would never be reached): refused, callers take the im2col route.
bool conv_tiles(int H, int W) {
  return (W % 128) == 0 || ((128 % W) == 0 && (((int64_t)H * W >= 128 && (H % (128 / W)) == 0) || ((int64_t)H * W < 128 && (128 % (H * W)) == 0)));
}
}  // namespace
}  // namespace o2345

extern "C" int o2345_conv3x3_f16(const void* x, int B, int H, int W, int C, const void* weight, int N, void* out, int64_t ldc,
                                 const o2345_epilogue* ep, float* splitk_ws, int64_t ws_floats, o2345_stream_t stream) {
  O2345_CHECK_ARG(x && weight && out, "null pointer");
  O2345_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && (C % 8) == 0 && N > 0, "bad sizes (C must be a multiple of 8)");
  O2345_CHECK_ARG(conv_tiles(H, W),
                  "implicit 3x3 conv: the image must tile into 128-pixel boxes (W % 128 == 0, or 128 % W == 0 with "
                  "H % (128 / W) == 0, or 128 % (H * W) == 0)");
  O2345_CHECK_ARG(((uintptr_t)x % 16) == 0 && ((uintptr_t)weight % 16) == 0, "operands must be 16-byte aligned");
  return conv_launch(x, B, H, W, C, weight, N, out, ldc, ep, splitk_ws, ws_floats, 9, 3, -1, -1, 0, 0, 0, (cudaStream_t)stream);
}

extern "C" int o2345_conv_up2x_f16(const void* x, int B, int H, int W, int C, const void* weight4, int N, void* out, int64_t ldc,
                                   const o2345_epilogue* ep, float* splitk_ws, int64_t ws_floats, o2345_stream_t stream) {
  O2345_CHECK_ARG(x && weight4 && out, "null pointer");
  O2345_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && (C % 8) == 0 && N > 0, "bad sizes (C must be a multiple of 8)");
  O2345_CHECK_ARG(conv_tiles(H, W), "up-sampling conv: the LOW-resolution image must tile into 128-pixel boxes (see o2345_conv3x3_f16)");
  O2345_CHECK_ARG(((uintptr_t)x % 16) == 0 && ((uintptr_t)weight4 % 16) == 0, "operands must be 16-byte aligned");
  O2345_CHECK_ARG(!ep || (!ep->residual && !ep->colstats && !ep->rowbias && ep->act != 3),
                  "up-sampling conv: bias / activation epilogues only (no residual, row bias, statistics, GEGLU)");
  O2345_CHECK_ARG((int64_t)B * 4 * H * W < (1ll << 31), "output rows must fit 31 bits");
  const __half* w = reinterpret_cast<const __half*>(weight4);
  for (int ph = 0; ph < 4; ++ph) {   // phase (a, b) = (row parity, column parity) of the output pixel
    const int a = ph >> 1, b = ph & 1;
    int rc = conv_launch(x, B, H, W, C, w + (int64_t)ph * N * 4 * C, N, out, ldc, ep, splitk_ws, ws_floats, 4, 2, b - 1, a - 1, 1, a, b,
                         (cudaStream_t)stream);
    if (rc) return rc;
  }
  return O2345_OK;
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
  p.ctaps = p.ctx = p.cox = p.coy = p.up = p.upa = p.upb = 0;
  set_workspace(p, splitk_ws, ws_floats);
  cudaStream_t st = (cudaStream_t)stream;
  const Config c = pick_config(p, cdiv(K, BK), nh == 0, ws_floats);
  p.splits = c.splits;
  CUtensorMap ma, mb;
  rc = make_map(&ma, A, M, K, lda, nh, nb, stride_a_h, stride_a_b, BM);
  if (rc) return rc;
  rc = make_map(&mb, B, N, K, ldb, nh, nb, stride_b_h, stride_b_b, c.bn / c.ctas);
  if (rc) return rc;
  return dispatch(c, pick_mode(p, c), ma, mb, p, nh > 0 ? nh * nb : 0, st);
}
