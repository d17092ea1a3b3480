// Cost-volume build: frustum mask, ordered compaction, fused back-project + variance/mean.
// Rows B3/B4/B5/B7 of SURVEY.md section 8.
//
//   frustum_mask_kernel   one thread per lattice voxel, all views: bit-exact restatement of
//                         reference ops/back_project.py:44-61 (separately rounded mul/add so
//                         that the CPU oracle reproduces every threshold decision);
//   compact_*             ascending-order stream compaction (kept-voxel order = ascending
//                         x*D*D + y*D + z, reference sparse_sdf_network.py:321-334);
//   costvol_gather_kernel four threads per kept voxel, each owning 4 of the 16 channels:
//                         per view one projection, four 16-byte taps from the channel-last
//                         feature map, running sum / sum of squares; the [Nv,V,16] tensor of
//                         the reference (1.76 GB at 96^3 x 32 views) is never materialised.
#include "common.cuh"

namespace o2345 {
namespace {

struct Proj {
  float gx, gy, z;
  bool vis;
};

// world -> normalised grid of one view; operation order matches oracle project_voxels().
__device__ __forceinline__ Proj project(const float* __restrict__ P, float wx, float wy, float wz,
                                        float size_w1, float size_h1) {
  float ix = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(P[0], wx), __fmul_rn(P[1], wy)), __fmul_rn(P[2], wz)), P[3]);
  float iy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(P[4], wx), __fmul_rn(P[5], wy)), __fmul_rn(P[6], wz)), P[7]);
  float iz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(P[8], wx), __fmul_rn(P[9], wy)), __fmul_rn(P[10], wz)), P[11]);
  if (iz >= 0.f) iz = fmaxf(iz, 1e-6f);
  float u = __fdiv_rn(ix, iz), v = __fdiv_rn(iy, iz);
  Proj r;
  r.gx = __fadd_rn(__fdiv_rn(__fmul_rn(2.f, u), size_w1), -1.f);
  r.gy = __fadd_rn(__fdiv_rn(__fmul_rn(2.f, v), size_h1), -1.f);
  r.z = iz;
  r.vis = (fabsf(r.gx) <= 1.f) && (fabsf(r.gy) <= 1.f) && (iz > 0.f);
  return r;
}

__device__ __forceinline__ void voxel_world(int64_t lin, int D, float vs, const float* __restrict__ origin,
                                            float& wx, float& wy, float& wz) {
  int z = (int)(lin % D);
  int y = (int)((lin / D) % D);
  int x = (int)(lin / ((int64_t)D * D));
  wx = __fadd_rn(__fmul_rn((float)x, vs), origin[0]);
  wy = __fadd_rn(__fmul_rn((float)y, vs), origin[1]);
  wz = __fadd_rn(__fmul_rn((float)z, vs), origin[2]);
}

__global__ void frustum_mask_kernel(const float* __restrict__ proj, int V, const float* __restrict__ origin,
                                    float vs, int D, float size_w1, float size_h1, int min_views,
                                    uint32_t* __restrict__ bits, uint8_t* __restrict__ keep) {
  extern __shared__ float sP[];  // [V][12]
  for (int i = threadIdx.x; i < V * 12; i += blockDim.x) sP[i] = proj[(i / 12) * 16 + (i % 12)];
  __syncthreads();
  int64_t n = (int64_t)D * D * D;
  int64_t lin = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lin >= n) return;
  float wx, wy, wz;
  voxel_world(lin, D, vs, origin, wx, wy, wz);
  uint32_t m = 0;
  for (int v = 0; v < V; ++v)
    if (project(sP + 12 * v, wx, wy, wz, size_w1, size_h1).vis) m |= (1u << v);
  bits[lin] = m;
  keep[lin] = __popc(m) > min_views ? 1 : 0;
}

// ---------------------------------------------------------------------------------------
// ordered compaction, 1024 elements per block
// ---------------------------------------------------------------------------------------
constexpr int CB = 1024;

__global__ void compact_count_kernel(const uint8_t* __restrict__ flags, int64_t n, int32_t* __restrict__ block_sums) {
  int64_t i = (int64_t)blockIdx.x * CB + threadIdx.x;
  int f = (i < n && flags[i]) ? 1 : 0;
  int c = __syncthreads_count(f);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = c;
}

// single block: exclusive scan of block_sums[nb] in place, total -> *count
__global__ void compact_scan_kernel(int32_t* __restrict__ block_sums, int nb, int32_t* __restrict__ count) {
  __shared__ int32_t warp_tot[32];
  __shared__ int32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += CB) {
    int i = base + threadIdx.x;
    int v = i < nb ? block_sums[i] : 0;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    if (lane == 31) warp_tot[w] = s;
    __syncthreads();
    if (w == 0) {
      int t = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += u;
      }
      warp_tot[lane] = t;
    }
    __syncthreads();
    int excl = s - v + (w > 0 ? warp_tot[w - 1] : 0) + carry;
    if (i < nb) block_sums[i] = excl;
    __syncthreads();
    if (threadIdx.x == CB - 1) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = carry;
}

__global__ void compact_scatter_kernel(const uint8_t* __restrict__ flags, int64_t n,
                                       const int32_t* __restrict__ block_offs, int32_t* __restrict__ rows,
                                       int32_t* __restrict__ index) {
  __shared__ int32_t warp_tot[32];
  int64_t i = (int64_t)blockIdx.x * CB + threadIdx.x;
  int f = (i < n && flags[i]) ? 1 : 0;
  unsigned b = __ballot_sync(0xffffffffu, f);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) warp_tot[w] = __popc(b);
  __syncthreads();
  if (w == 0) {
    int t = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int u = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += u;
    }
    warp_tot[lane] = t;
  }
  __syncthreads();
  int pos = block_offs[blockIdx.x] + (w > 0 ? warp_tot[w - 1] : 0) + __popc(b & ((1u << lane) - 1u));
  if (i < n) {
    if (f) rows[pos] = (int32_t)i;
    if (index) index[i] = f ? pos : -1;
  }
}

// ---------------------------------------------------------------------------------------
// fused back-projection + variance/mean
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
costvol_gather_kernel(const float* __restrict__ feats, int V, int h, int w, float size_w1, float size_h1,
                      const float* __restrict__ proj, const float* __restrict__ origin, float vs, int D,
                      const int32_t* __restrict__ rows, const int32_t* __restrict__ count,
                      const uint32_t* __restrict__ bits, float* __restrict__ cost) {
  extern __shared__ float sP[];
  for (int i = threadIdx.x; i < V * 12; i += blockDim.x) sP[i] = proj[(i / 12) * 16 + (i % 12)];
  __syncthreads();
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t row = t >> 2;
  int q = (int)(t & 3);  // channel quad
  if (row >= *count) return;
  int64_t lin = rows[row];
  float wx, wy, wz;
  voxel_world(lin, D, vs, origin, wx, wy, wz);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), sq = s;
  const float wm1 = (float)(w - 1), hm1 = (float)(h - 1);
  for (int v = 0; v < V; ++v) {
    Proj p = project(sP + 12 * v, wx, wy, wz, size_w1, size_h1);
    // ATen grid_sampler_2d, bilinear, zeros padding, align_corners=True
    float fx = ((p.gx + 1.f) / 2.f) * wm1, fy = ((p.gy + 1.f) / 2.f) * hm1;
    float x0 = floorf(fx), y0 = floorf(fy);
    // all four taps outside the map (or non-finite coordinates): the view contributes zeros
    if (!(x0 >= -1.f && x0 <= wm1 && y0 >= -1.f && y0 <= hm1)) continue;
    float x1 = x0 + 1.f, y1 = y0 + 1.f;
    float wnw = (x1 - fx) * (y1 - fy), wne = (fx - x0) * (y1 - fy);
    float wsw = (x1 - fx) * (fy - y0), wse = (fx - x0) * (fy - y0);
    int ix0 = (int)x0, iy0 = (int)y0;
    bool inx0 = ix0 >= 0, inx1 = ix0 + 1 <= w - 1, iny0 = iy0 >= 0, iny1 = iy0 + 1 <= h - 1;
    const float* base = feats + (((int64_t)v * h + iy0) * w + ix0) * 16 + 4 * q;
    float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
    if (iny0 && inx0) { float4 a = ldg4(base); f.x = fmaf(a.x, wnw, f.x); f.y = fmaf(a.y, wnw, f.y); f.z = fmaf(a.z, wnw, f.z); f.w = fmaf(a.w, wnw, f.w); }
    if (iny0 && inx1) { float4 a = ldg4(base + 16); f.x = fmaf(a.x, wne, f.x); f.y = fmaf(a.y, wne, f.y); f.z = fmaf(a.z, wne, f.z); f.w = fmaf(a.w, wne, f.w); }
    if (iny1 && inx0) { float4 a = ldg4(base + (int64_t)w * 16); f.x = fmaf(a.x, wsw, f.x); f.y = fmaf(a.y, wsw, f.y); f.z = fmaf(a.z, wsw, f.z); f.w = fmaf(a.w, wsw, f.w); }
    if (iny1 && inx1) { float4 a = ldg4(base + (int64_t)w * 16 + 16); f.x = fmaf(a.x, wse, f.x); f.y = fmaf(a.y, wse, f.y); f.z = fmaf(a.z, wse, f.z); f.w = fmaf(a.w, wse, f.w); }
    s.x += f.x; s.y += f.y; s.z += f.z; s.w += f.w;
    sq.x = fmaf(f.x, f.x, sq.x); sq.y = fmaf(f.y, f.y, sq.y); sq.z = fmaf(f.z, f.z, sq.z); sq.w = fmaf(f.w, f.w, sq.w);
  }
  float inv = 1.f / ((float)__popc(bits[lin]) + 1e-5f);
  float4 mean = make_float4(s.x * inv, s.y * inv, s.z * inv, s.w * inv);
  float4 var = make_float4(sq.x * inv - mean.x * mean.x, sq.y * inv - mean.y * mean.y,
                           sq.z * inv - mean.z * mean.z, sq.w * inv - mean.w * mean.w);
  *reinterpret_cast<float4*>(cost + row * 32 + 4 * q) = var;
  *reinterpret_cast<float4*>(cost + row * 32 + 16 + 4 * q) = mean;
}

__global__ void dense_scatter_kernel(const float* __restrict__ feat, const int32_t* __restrict__ rows,
                                     const int32_t* __restrict__ count, int64_t n_cells,
                                     float* __restrict__ vol_cl, float* __restrict__ vol_cf, float* __restrict__ occ) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t row = t >> 2;
  int q = (int)(t & 3);
  if (row >= *count) return;
  int64_t lin = rows[row];
  float4 v = ldg4(feat + row * 16 + 4 * q);
  *reinterpret_cast<float4*>(vol_cl + lin * 16 + 4 * q) = v;
  if (vol_cf) {
    vol_cf[(int64_t)(4 * q + 0) * n_cells + lin] = v.x;
    vol_cf[(int64_t)(4 * q + 1) * n_cells + lin] = v.y;
    vol_cf[(int64_t)(4 * q + 2) * n_cells + lin] = v.z;
    vol_cf[(int64_t)(4 * q + 3) * n_cells + lin] = v.w;
  }
  if (q == 0) occ[lin] = 1.f;
}

// occupancy lookup: ATen grid_sample(mode='nearest', align_corners=False) after the xyz->zyx flip
// (reference sparse_neus_renderer.py:153-169); idx = nearbyint(((p+1)*D-1)/2) per axis.
__global__ void occ_nearest_kernel(o2345_points src, int64_t n, const float* __restrict__ occ, int D,
                                   uint8_t* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float p[3];
  if (src.mode == O2345_PTS_EXPLICIT) {
    p[0] = src.pts[3 * i], p[1] = src.pts[3 * i + 1], p[2] = src.pts[3 * i + 2];
  } else {
    int64_t r = i / src.S;
    int s = (int)(i - r * src.S);
    float t = src.z[r * src.z_stride + s];
    for (int a = 0; a < 3; ++a) p[a] = __fadd_rn(src.rays_o[3 * r + a], __fmul_rn(src.rays_d[3 * r + a], t));
  }
  int idx[3];
  bool ok = true;
  for (int a = 0; a < 3; ++a) {
    float f = nearbyintf(((p[a] + 1.f) * (float)D - 1.f) / 2.f);
    ok = ok && (f >= 0.f) && (f <= (float)(D - 1));
    idx[a] = (int)fminf(fmaxf(f, 0.f), (float)(D - 1));
  }
  float v = ok ? occ[((int64_t)idx[0] * D + idx[1]) * D + idx[2]] : 0.f;
  out[i] = v > 0.f ? 1 : 0;
}

}  // namespace
}  // namespace o2345

using namespace o2345;

extern "C" int o2345_frustum_mask(const float* proj, int V, const float* origin, float voxel_size, int D,
                                  int sizeH, int sizeW, int min_views, uint32_t* mask_bits, uint8_t* keep,
                                  o2345_stream_t stream) {
  O2345_CHECK_ARG(proj && origin && mask_bits && keep, "null pointer");
  O2345_CHECK_ARG(V >= 1 && V <= 32, "1..32 views supported (mask is one 32-bit word per voxel)");
  O2345_CHECK_ARG(D >= 2 && D <= 1024 && sizeH > 1 && sizeW > 1, "bad sizes");
  int64_t n = (int64_t)D * D * D;
  frustum_mask_kernel<<<cdiv(n, 256), 256, V * 12 * sizeof(float), (cudaStream_t)stream>>>(
      proj, V, origin, voxel_size, D, (float)(sizeW - 1), (float)(sizeH - 1), min_views, mask_bits, keep);
  O2345_LAUNCH_CHECK();
  return O2345_OK;
}

extern "C" int64_t o2345_compact_scratch_ints(int64_t n) { return (n + CB - 1) / CB + 1; }

extern "C" int o2345_compact(const uint8_t* flags, int64_t n, int32_t* rows, int32_t* index, int32_t* count,
                             int32_t* scratch, o2345_stream_t stream) {
  O2345_CHECK_ARG(flags && rows && count && scratch, "null pointer");
  O2345_CHECK_ARG(n > 0 && n < ((int64_t)1 << 31), "element count out of range");
  int nb = cdiv(n, CB);
  cudaStream_t st = (cudaStream_t)stream;
  compact_count_kernel<<<nb, CB, 0, st>>>(flags, n, scratch);
  compact_scan_kernel<<<1, CB, 0, st>>>(scratch, nb, count);
  compact_scatter_kernel<<<nb, CB, 0, st>>>(flags, n, scratch, rows, index);
  O2345_LAUNCH_CHECK();
  return O2345_OK;
}

extern "C" int o2345_costvol_gather(const float* feats_nhwc, int V, int h, int w, int sizeH, int sizeW,
                                    const float* proj, const float* origin, float voxel_size, int D,
                                    const int32_t* rows, const int32_t* count, int64_t max_rows,
                                    const uint32_t* mask_bits, float* cost, o2345_stream_t stream) {
  O2345_CHECK_ARG(feats_nhwc && proj && origin && rows && count && mask_bits && cost, "null pointer");
  O2345_CHECK_ARG(V >= 1 && V <= 32 && h > 1 && w > 1 && sizeH > 1 && sizeW > 1 && max_rows > 0, "bad sizes");
  costvol_gather_kernel<<<cdiv(max_rows * 4, 256), 256, V * 12 * sizeof(float), (cudaStream_t)stream>>>(
      feats_nhwc, V, h, w, (float)(sizeW - 1), (float)(sizeH - 1), proj, origin, voxel_size, D, rows, count,
      mask_bits, cost);
  O2345_LAUNCH_CHECK();
  return O2345_OK;
}

extern "C" int o2345_dense_scatter(const float* feat, const int32_t* rows, const int32_t* count, int64_t max_rows,
                                   int D, float* vol_cl, float* vol_cf, float* occ, o2345_stream_t stream) {
  O2345_CHECK_ARG(feat && rows && count && vol_cl && occ, "null pointer");
  int64_t n = (int64_t)D * D * D;
  cudaStream_t st = (cudaStream_t)stream;
  O2345_CUDA(cudaMemsetAsync(vol_cl, 0, n * 16 * sizeof(float), st));
  if (vol_cf) O2345_CUDA(cudaMemsetAsync(vol_cf, 0, n * 16 * sizeof(float), st));
  O2345_CUDA(cudaMemsetAsync(occ, 0, n * sizeof(float), st));
  dense_scatter_kernel<<<cdiv(max_rows * 4, 256), 256, 0, st>>>(feat, rows, count, n, vol_cl, vol_cf, occ);
  O2345_LAUNCH_CHECK();
  return O2345_OK;
}

extern "C" int o2345_occ_nearest(const o2345_points* src, int64_t n, const float* occ, int D, uint8_t* out,
                                 o2345_stream_t stream) {
  O2345_CHECK_ARG(src && occ && out, "null pointer");
  O2345_CHECK_ARG(src->mode == O2345_PTS_EXPLICIT || src->mode == O2345_PTS_RAYS, "explicit or ray points only");
  if (n == 0) return O2345_OK;
  occ_nearest_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(*src, n, occ, D, out);
  O2345_LAUNCH_CHECK();
  return O2345_OK;
}
