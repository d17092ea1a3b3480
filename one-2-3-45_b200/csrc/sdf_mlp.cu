// SDF query kernel: quirky trilinear latent fetch + positional embedding + weight-normed MLP,
// with an optional analytic gradient (reverse mode through the three layers, the embedding
// and the trilinear weights).  Rows B8/B9/B10 of SURVEY.md section 8.
//
// One persistent CTA per SM (grid = #SMs).  A CTA processes tiles of 128 query points:
//   stage 0  gather the 16-channel latent (8 corners x 64 B, channel-last volume) and build
//            the 39-wide positional embedding; both land in shared memory k-major
//            ([feature][point]) so that the GEMM reads them as conflict-free float4s;
//   stage 1-3  three register-tiled fp32 GEMMs (8 points x 8 outputs per thread) against the
//            layer weights, which are streamed L2 -> shared memory once per tile and layer;
//   stage 4  (grad only) two transposed GEMMs for the backward pass and a per-point
//            contraction with d(embedding)/dx and d(trilinear weights)/dx.
// Arithmetic is fp32 FMA throughout (the reference is fp32; tolerance is stated in the tests).
#include "common.cuh"

namespace o2345 {
namespace {

constexpr int TM = 128;   // points per tile
constexpr int NT = 256;   // threads per CTA
constexpr int PE = O2345_SDF_PE, HID = O2345_SDF_HID, LAT = O2345_SDF_LAT, IN1 = O2345_SDF_IN1;
constexpr int W0PAD = 48;  // W0 rows padded to 48 columns for the backward GEMM

constexpr int OFF_W0T = 0;
constexpr int OFF_B0 = OFF_W0T + PE * HID;
constexpr int OFF_W1T = OFF_B0 + HID;
constexpr int OFF_B1 = OFF_W1T + IN1 * HID;
constexpr int OFF_W2T = OFF_B1 + HID;
constexpr int OFF_B2 = OFF_W2T + IN1 * HID;
constexpr int OFF_W1 = OFF_B2 + HID;          // [128][144]
constexpr int OFF_W0 = OFF_W1 + HID * IN1;    // [128][48]
constexpr int PACK_FLOATS = OFF_W0 + HID * W0PAD;

constexpr int SM_ACT = IN1 * TM;              // 18432 floats
constexpr int SM_W = IN1 * HID;               // 18432 floats (also output staging [TM][129])
constexpr int SM_A0 = HID * TM;               // 16384 floats (grad only)
constexpr int SM_MISC = 8 * TM;               // pts(3) + grad partials(3) + flags
constexpr int SMEM_FWD = (SM_ACT + SM_W + SM_MISC) * 4;
constexpr int SMEM_GRAD = (SM_ACT + SM_W + SM_A0 + SM_MISC) * 4;

__device__ __forceinline__ float softplus100(float x) {
  float bx = 100.f * x;
  return bx > 20.f ? x : log1pf(expf(bx)) * 0.01f;
}
// d softplus / dx expressed through the activation a = softplus(x): sigmoid(100x) = 1 - exp(-100a)
__device__ __forceinline__ float dsoftplus_from_act(float a) { return -expm1f(-100.f * a); }

struct Tri {
  int base[3];     // clamped floor index per axis (x,y,z)
  int hi[3];       // clamped floor+1 index
  float w0[3], w1[3];
  bool inb;
};

// reference ops/grid_sampler.py:79-90 (after the xyz->zyx flip of sparse_sdf_network.py:408)
__device__ __forceinline__ Tri tri_setup(float px, float py, float pz, int D) {
  Tri t;
  float p[3] = {px, py, pz};
  bool inb = true;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float tt = __fmul_rn(__fdiv_rn(__fadd_rn(p[a], 1.f), 2.f), (float)(D - 1));
    inb = inb && (tt > 0.f) && (tt < (float)D);
    float f = floorf(tt);
    t.w1[a] = tt - f;
    t.w0[a] = (f + 1.f) - tt;
    // clamp in float first: tt may be huge or NaN for far-away points
    float fl = fminf(fmaxf(f, 0.f), (float)(D - 1));
    float fh = fminf(fmaxf(f + 1.f, 0.f), (float)(D - 1));
    t.base[a] = (int)fl;
    t.hi[a] = (int)fh;
  }
  t.inb = inb;
  return t;
}

__device__ __forceinline__ void load_point(const o2345_points& src, int64_t gi, float& x, float& y, float& z) {
  if (src.mode == O2345_PTS_EXPLICIT) {
    x = __ldg(src.pts + 3 * gi), y = __ldg(src.pts + 3 * gi + 1), z = __ldg(src.pts + 3 * gi + 2);
  } else if (src.mode == O2345_PTS_LATTICE) {
    int R = src.R;
    int64_t ix = gi / ((int64_t)R * R);
    int iy = (int)((gi / R) % R), iz = (int)(gi % R);
    x = __ldg(src.lin + ix), y = __ldg(src.lin + iy), z = __ldg(src.lin + iz);
  } else {
    int64_t r = gi / src.S;
    int s = (int)(gi - r * src.S);
    float t = __ldg(src.z + r * src.z_stride + s);
    // o + d * t with separately rounded multiply and add (torch evaluates it that way)
    x = __fadd_rn(__ldg(src.rays_o + 3 * r), __fmul_rn(__ldg(src.rays_d + 3 * r), t));
    y = __fadd_rn(__ldg(src.rays_o + 3 * r + 1), __fmul_rn(__ldg(src.rays_d + 3 * r + 1), t));
    z = __fadd_rn(__ldg(src.rays_o + 3 * r + 2), __fmul_rn(__ldg(src.rays_d + 3 * r + 2), t));
  }
}

// acc[i][j] += sum_k A[k][m0+i] * B[k][n0+j];  A k-major with row stride TM, B row stride 128.
template <int K>
__device__ __forceinline__ void gemm_fwd(const float* __restrict__ sA, const float* __restrict__ sB,
                                         float (&acc)[8][8], int m0, int n0) {
#pragma unroll 2
  for (int k = 0; k < K; ++k) {
    float4 a0 = *reinterpret_cast<const float4*>(sA + k * TM + m0);
    float4 a1 = *reinterpret_cast<const float4*>(sA + k * TM + m0 + 4);
    float4 b0 = *reinterpret_cast<const float4*>(sB + k * HID + n0);
    float4 b1 = *reinterpret_cast<const float4*>(sB + k * HID + n0 + 4);
    float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

// acc[i][j] += sum_{k<128} A[k][m0+i] * B[k][ty + 16 j];  B row stride LDB, NJ columns per thread.
template <int NJ, int LDB>
__device__ __forceinline__ void gemm_bwd(const float* __restrict__ sA, const float* __restrict__ sB,
                                         float (&acc)[8][NJ], int m0, int ty) {
#pragma unroll 2
  for (int k = 0; k < HID; ++k) {
    float4 a0 = *reinterpret_cast<const float4*>(sA + k * TM + m0);
    float4 a1 = *reinterpret_cast<const float4*>(sA + k * TM + m0 + 4);
    float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    float b[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) b[j] = sB[k * LDB + ty + 16 * j];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

__device__ __forceinline__ void load_weights(float* sW, const float* __restrict__ g, int nfloats) {
  for (int i = threadIdx.x * 4; i < nfloats; i += NT * 4)
    *reinterpret_cast<float4*>(sW + i) = ldg4(g + i);
}

template <bool GRAD>
__global__ void __launch_bounds__(NT, 1)
sdf_query_kernel(o2345_points src, int64_t n, const float* __restrict__ vol, int D,
                 const float* __restrict__ wp, const uint8_t* __restrict__ active, float inactive_sdf,
                 float sign, float* __restrict__ o_sdf, float* __restrict__ o_feat,
                 float* __restrict__ o_lat, float* __restrict__ o_grad) {
  extern __shared__ __align__(16) float smem[];
  float* sAct = smem;                       // [144][TM]
  float* sW = sAct + SM_ACT;                // weights of the current layer / output staging
  float* sA0 = sW + SM_W;                   // [128][TM] layer-0 activations (GRAD only)
  float* sMisc = GRAD ? sA0 + SM_A0 : sA0;  // pts[3][TM], gpart[3][TM], flag[TM]
  float* sPts = sMisc;
  float* sGp = sMisc + 3 * TM;
  int* sFlag = reinterpret_cast<int*>(sMisc + 6 * TM);

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = tx * 8, n0 = ty * 8;
  const int pm = tid & (TM - 1), half = tid >> 7;
  const int64_t ntiles = (n + TM - 1) / TM;

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t g0 = tile * TM;
    const int64_t gi = g0 + pm;
    // ---------------- stage 0: point, latent, embedding --------------------------------
    bool act = gi < n && (active == nullptr || active[gi] != 0);
    int any = __syncthreads_or(act ? 1 : 0);
    if (!any) {  // nothing to evaluate in this tile: defaults only (block-uniform branch)
      if (half == 0 && gi < n) {
        if (o_sdf) o_sdf[gi] = inactive_sdf;
        if (o_grad) { o_grad[3 * gi] = 0.f; o_grad[3 * gi + 1] = 0.f; o_grad[3 * gi + 2] = 0.f; }
      }
      int64_t cnt = min((int64_t)TM, n - g0);
      if (o_feat) for (int64_t e = tid; e < cnt * 127; e += NT) o_feat[g0 * 127 + e] = 0.f;
      if (o_lat) for (int64_t e = tid; e < cnt * LAT; e += NT) o_lat[g0 * LAT + e] = 0.f;
      continue;
    }
    float px = 0.f, py = 0.f, pz = 0.f;
    if (act) load_point(src, gi, px, py, pz);
    if (half == 0) {
      sPts[pm] = px, sPts[TM + pm] = py, sPts[2 * TM + pm] = pz;
      sFlag[pm] = act ? 1 : 0;
    }
    {
      // latent channels [8*half, 8*half+8)
      Tri t = tri_setup(px, py, pz, D);
      float lat[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) lat[c] = 0.f;
      if (act && t.inb) {
#pragma unroll
        for (int corner = 0; corner < 8; ++corner) {
          int dx = corner >> 2, dy = (corner >> 1) & 1, dz = corner & 1;
          int ix = dx ? t.hi[0] : t.base[0], iy = dy ? t.hi[1] : t.base[1], iz = dz ? t.hi[2] : t.base[2];
          // reference weight product order: (w_Z * w_Y) * w_X  (its ix is our z axis)
          float w = ((dz ? t.w1[2] : t.w0[2]) * (dy ? t.w1[1] : t.w0[1])) * (dx ? t.w1[0] : t.w0[0]);
          const float* v = vol + (((int64_t)ix * D + iy) * D + iz) * LAT + 8 * half;
          float4 v0 = ldg4(v), v1 = ldg4(v + 4);
          lat[0] = fmaf(v0.x, w, lat[0]); lat[1] = fmaf(v0.y, w, lat[1]);
          lat[2] = fmaf(v0.z, w, lat[2]); lat[3] = fmaf(v0.w, w, lat[3]);
          lat[4] = fmaf(v1.x, w, lat[4]); lat[5] = fmaf(v1.y, w, lat[5]);
          lat[6] = fmaf(v1.z, w, lat[6]); lat[7] = fmaf(v1.w, w, lat[7]);
        }
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) sAct[(HID + 8 * half + c) * TM + pm] = lat[c];
      // embedding: half 0 -> raw xyz + frequencies 0..2, half 1 -> frequencies 3..5
      float p[3] = {px, py, pz};
      if (half == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) sAct[a * TM + pm] = p[a];
      }
#pragma unroll
      for (int f = 0; f < 3; ++f) {
        int k = 3 * half + f;
        float fr = (float)(1 << k);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          float s, c;
          sincosf(fr * p[a], &s, &c);
          sAct[(3 + 6 * k + a) * TM + pm] = s;
          sAct[(3 + 6 * k + 3 + a) * TM + pm] = c;
        }
      }
    }
    load_weights(sW, wp + OFF_W0T, PE * HID);
    __syncthreads();

    float acc[8][8];
    // ---------------- layer 0: 39 -> 128, softplus ------------------------------------
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float b = __ldg(wp + OFF_B0 + n0 + j);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i][j] = b;
    }
    gemm_fwd<PE>(sAct, sW, acc, m0, n0);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 v0 = make_float4(softplus100(acc[0][j]), softplus100(acc[1][j]), softplus100(acc[2][j]), softplus100(acc[3][j]));
      float4 v1 = make_float4(softplus100(acc[4][j]), softplus100(acc[5][j]), softplus100(acc[6][j]), softplus100(acc[7][j]));
      *reinterpret_cast<float4*>(sAct + (n0 + j) * TM + m0) = v0;
      *reinterpret_cast<float4*>(sAct + (n0 + j) * TM + m0 + 4) = v1;
      if (GRAD) {
        *reinterpret_cast<float4*>(sA0 + (n0 + j) * TM + m0) = v0;
        *reinterpret_cast<float4*>(sA0 + (n0 + j) * TM + m0 + 4) = v1;
      }
    }
    load_weights(sW, wp + OFF_W1T, IN1 * HID);
    __syncthreads();
    // ---------------- layer 1: 144 -> 128, softplus -----------------------------------
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float b = __ldg(wp + OFF_B1 + n0 + j);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i][j] = b;
    }
    gemm_fwd<IN1>(sAct, sW, acc, m0, n0);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      *reinterpret_cast<float4*>(sAct + (n0 + j) * TM + m0) =
          make_float4(softplus100(acc[0][j]), softplus100(acc[1][j]), softplus100(acc[2][j]), softplus100(acc[3][j]));
      *reinterpret_cast<float4*>(sAct + (n0 + j) * TM + m0 + 4) =
          make_float4(softplus100(acc[4][j]), softplus100(acc[5][j]), softplus100(acc[6][j]), softplus100(acc[7][j]));
    }
    load_weights(sW, wp + OFF_W2T, IN1 * HID);
    __syncthreads();
    // ---------------- layer 2: 144 -> 128 (no activation) -----------------------------
    const bool need_feat = (o_feat != nullptr);
    if (need_feat) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float b = __ldg(wp + OFF_B2 + n0 + j);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i][j] = b;
      }
      gemm_fwd<IN1>(sAct, sW, acc, m0, n0);
      __syncthreads();  // everyone is done reading sW -> reuse it as staging [TM][129]
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) sW[(m0 + i) * 129 + n0 + j] = acc[i][j];
      __syncthreads();
      int64_t cnt = min((int64_t)TM, n - g0);
      for (int64_t e = tid; e < cnt * 127; e += NT) {
        int m = (int)(e / 127), c = (int)(e - (int64_t)m * 127);
        o_feat[g0 * 127 + e] = sFlag[m] ? sW[m * 129 + 1 + c] : 0.f;
      }
      if (half == 0 && gi < n && o_sdf) o_sdf[gi] = sFlag[pm] ? sign * sW[pm * 129] : inactive_sdf;
    } else {
      // only the sdf column is needed: one dot product per point (K=144) by threads 0..127
      if (half == 0) {
        float s = __ldg(wp + OFF_B2);
#pragma unroll 8
        for (int k = 0; k < IN1; ++k) s = fmaf(sAct[k * TM + pm], sW[k * HID], s);
        if (gi < n && o_sdf) o_sdf[gi] = sFlag[pm] ? sign * s : inactive_sdf;
      }
    }
    if (o_lat) {
      int64_t cnt = min((int64_t)TM, n - g0);
      for (int64_t e = tid; e < cnt * LAT; e += NT) {
        int m = (int)(e >> 4), c = (int)(e & 15);
        o_lat[g0 * LAT + e] = sFlag[m] ? sAct[(HID + c) * TM + m] : 0.f;
      }
    }
    if (GRAD) {
      // ---------------- backward: d sdf / d (layer-2 input) = row 0 of W2 --------------
      // sW currently holds W2t [144][128] (or the staging copy if feat was written, so read
      // row 0 of W2 from global instead: W2t[k][0]).
      __syncthreads();
      // delta1[j][m] = W2[0][j] * softplus'(z1) ;  a1 lives in sAct rows 0..127
      for (int e = tid; e < HID * TM; e += NT) {
        int j = e >> 7;
        float w = __ldg(wp + OFF_W2T + j * HID);
        sAct[e] = w * dsoftplus_from_act(sAct[e]);
      }
      load_weights(sW, wp + OFF_W1, HID * IN1);
      __syncthreads();
      {
        float g[8][9];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 9; ++j) g[i][j] = 0.f;
        gemm_bwd<9, IN1>(sAct, sW, g, m0, ty);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          int nn = ty + 16 * j;
          float v[8];
          if (j < 8) {  // delta0 = g * softplus'(z0)
            float4 a0 = *reinterpret_cast<const float4*>(sA0 + nn * TM + m0);
            float4 a1 = *reinterpret_cast<const float4*>(sA0 + nn * TM + m0 + 4);
            float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = g[i][j] * dsoftplus_from_act(a[i]);
          } else {      // gradient w.r.t. the latent: direct path through layer 2 + layer 1
            float w = __ldg(wp + OFF_W2T + nn * HID);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = g[i][j] + w;
          }
          *reinterpret_cast<float4*>(sAct + nn * TM + m0) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(sAct + nn * TM + m0 + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
      }
      load_weights(sW, wp + OFF_W0, HID * W0PAD);
      __syncthreads();
      {
        float g[8][3];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) g[i][j] = 0.f;
        gemm_bwd<3, W0PAD>(sAct, sW, g, m0, ty);
        // sA0 is free now (delta0 already formed): g_pe[n][m] -> sA0 rows 0..47
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          int nn = ty + 16 * j;
          *reinterpret_cast<float4*>(sA0 + nn * TM + m0) = make_float4(g[0][j], g[1][j], g[2][j], g[3][j]);
          *reinterpret_cast<float4*>(sA0 + nn * TM + m0 + 4) = make_float4(g[4][j], g[5][j], g[6][j], g[7][j]);
        }
      }
      __syncthreads();
      // ---------------- per point: embedding part (half 0) + trilinear part (half 1) ---
      {
        float qx = sPts[pm], qy = sPts[TM + pm], qz = sPts[2 * TM + pm];
        float gx = 0.f, gy = 0.f, gz = 0.f;
        if (half == 0) {
          float p[3] = {qx, qy, qz};
          float gg[3] = {sA0[0 * TM + pm], sA0[1 * TM + pm], sA0[2 * TM + pm]};
#pragma unroll
          for (int k = 0; k < 6; ++k) {
            float fr = (float)(1 << k);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
              float s, c;
              sincosf(fr * p[a], &s, &c);
              float gs = sA0[(3 + 6 * k + a) * TM + pm], gc = sA0[(3 + 6 * k + 3 + a) * TM + pm];
              gg[a] = fmaf(fr, gs * c - gc * s, gg[a]);
            }
          }
          gx = gg[0], gy = gg[1], gz = gg[2];
        } else {
          Tri t = tri_setup(qx, qy, qz, D);
          if (t.inb) {
            float gl[LAT];
#pragma unroll
            for (int c = 0; c < LAT; ++c) gl[c] = sAct[(HID + c) * TM + pm];
            float sc = 0.5f * (float)(D - 1);  // d t / d p
#pragma unroll
            for (int corner = 0; corner < 8; ++corner) {
              int dx = corner >> 2, dy = (corner >> 1) & 1, dz = corner & 1;
              int ix = dx ? t.hi[0] : t.base[0], iy = dy ? t.hi[1] : t.base[1], iz = dz ? t.hi[2] : t.base[2];
              const float* v = vol + (((int64_t)ix * D + iy) * D + iz) * LAT;
              float dot = 0.f;
#pragma unroll
              for (int c4 = 0; c4 < 4; ++c4) {
                float4 vv = ldg4(v + 4 * c4);
                dot = fmaf(vv.x, gl[4 * c4], dot); dot = fmaf(vv.y, gl[4 * c4 + 1], dot);
                dot = fmaf(vv.z, gl[4 * c4 + 2], dot); dot = fmaf(vv.w, gl[4 * c4 + 3], dot);
              }
              float wx = dx ? t.w1[0] : t.w0[0], wy = dy ? t.w1[1] : t.w0[1], wz = dz ? t.w1[2] : t.w0[2];
              float sx = dx ? sc : -sc, sy = dy ? sc : -sc, sz = dz ? sc : -sc;
              gx = fmaf(dot, sx * wy * wz, gx);
              gy = fmaf(dot, wx * sy * wz, gy);
              gz = fmaf(dot, wx * wy * sz, gz);
            }
          }
          sGp[pm] = gx, sGp[TM + pm] = gy, sGp[2 * TM + pm] = gz;
        }
        __syncthreads();
        if (half == 0 && gi < n && o_grad) {
          bool on = sFlag[pm] != 0;
          o_grad[3 * gi] = on ? gx + sGp[pm] : 0.f;
          o_grad[3 * gi + 1] = on ? gy + sGp[TM + pm] : 0.f;
          o_grad[3 * gi + 2] = on ? gz + sGp[2 * TM + pm] : 0.f;
        }
      }
    }
    __syncthreads();  // smem is reused by the next tile
  }
}

__global__ void pack_weights_kernel(const float* w0, const float* b0, const float* w1, const float* b1,
                                    const float* w2, const float* b2, float* pack) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= PACK_FLOATS) return;
  float v;
  if (i < OFF_B0) { int k = i / HID, o = i % HID; v = w0[o * PE + k]; }
  else if (i < OFF_W1T) v = b0[i - OFF_B0];
  else if (i < OFF_B1) { int e = i - OFF_W1T; int k = e / HID, o = e % HID; v = w1[o * IN1 + k]; }
  else if (i < OFF_W2T) v = b1[i - OFF_B1];
  else if (i < OFF_B2) { int e = i - OFF_W2T; int k = e / HID, o = e % HID; v = w2[o * IN1 + k]; }
  else if (i < OFF_W1) v = b2[i - OFF_B2];
  else if (i < OFF_W0) v = w1[i - OFF_W1];
  else { int e = i - OFF_W0; int o = e / W0PAD, k = e % W0PAD; v = k < PE ? w0[o * PE + k] : 0.f; }
  pack[i] = v;
}

}  // namespace
}  // namespace o2345

using namespace o2345;

static_assert(PACK_FLOATS == O2345_SDF_PACK_FLOATS, "header and kernel disagree on the pack size");

extern "C" int o2345_sdf_pack_weights(const float* w0, const float* b0, const float* w1, const float* b1,
                                      const float* w2, const float* b2, float* pack, o2345_stream_t stream) {
  O2345_CHECK_ARG(w0 && b0 && w1 && b1 && w2 && b2 && pack, "null pointer");
  pack_weights_kernel<<<cdiv(PACK_FLOATS, 256), 256, 0, (cudaStream_t)stream>>>(w0, b0, w1, b1, w2, b2, pack);
  O2345_LAUNCH_CHECK();
  return O2345_OK;
}

extern "C" int o2345_sdf_query(const o2345_points* src, int64_t n, const float* vol_cl, int D,
                               const float* wpack, const uint8_t* active, float inactive_sdf, int negate,
                               float* sdf, float* feat, float* latent, float* grad, o2345_stream_t stream) {
  O2345_CHECK_ARG(n >= 0, "negative point count");
  if (n == 0) return O2345_OK;
  O2345_CHECK_ARG(src && vol_cl && wpack, "null pointer");
  O2345_CHECK_ARG(D >= 2 && D <= 1024, "volume side out of range");
  if (src->mode == O2345_PTS_EXPLICIT) O2345_CHECK_ARG(src->pts, "explicit points missing");
  else if (src->mode == O2345_PTS_LATTICE) O2345_CHECK_ARG(src->lin && src->R > 0 && n == (int64_t)src->R * src->R * src->R, "bad lattice");
  else if (src->mode == O2345_PTS_RAYS) O2345_CHECK_ARG(src->rays_o && src->rays_d && src->z && src->S > 0 && src->z_stride >= src->S && n % src->S == 0, "bad ray source");
  else O2345_CHECK_ARG(false, "unknown point source mode");
  static bool attr_done = false;
  if (!attr_done) {
    O2345_CUDA(cudaFuncSetAttribute(sdf_query_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_FWD));
    O2345_CUDA(cudaFuncSetAttribute(sdf_query_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_GRAD));
    attr_done = true;
  }
  int64_t ntiles = (n + TM - 1) / TM;
  int grid = (int)(ntiles < (int64_t)sm_count() ? ntiles : (int64_t)sm_count());
  float sign = negate ? -1.f : 1.f;
  if (grad)
    sdf_query_kernel<true><<<grid, NT, SMEM_GRAD, (cudaStream_t)stream>>>(*src, n, vol_cl, D, wpack, active, inactive_sdf, sign, sdf, feat, latent, grad);
  else
    sdf_query_kernel<false><<<grid, NT, SMEM_FWD, (cudaStream_t)stream>>>(*src, n, vol_cl, D, wpack, active, inactive_sdf, sign, sdf, feat, latent, grad);
  O2345_LAUNCH_CHECK();
  return O2345_OK;
}
