/*
 * libo2345_sm100.so -- C-ABI of the B200-native (sm_100a) kernels behind One-2-3-45's
 * SparseNeuS-style reconstruction hot path.
 *
 * The reference has no FFI of its own for this path: its boundary is plain Python classes
 * (SparseSdfNetwork, SparseNeuSRenderer, FeatureNet, GeneralRenderingNetwork, Projector) that
 * call PyTorch/ATen, torchsparse v1.4.0, inplace_abn and PyMCubes.  Each entry point below
 * names the reference call site(s) it replaces; INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - all floating point data is fp32, dense and contiguous in the stated layout;
 *   - the caller allocates every buffer (outputs and scratch); nothing is allocated or
 *     freed behind the ABI and no call synchronises the device;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it;
 *   - return value 0 on success, negative O2345_E* otherwise; o2345_last_error() returns a
 *     thread-local description of the most recent failure.
 */
#ifndef O2345_H_
#define O2345_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define O2345_OK 0
#define O2345_EINVAL (-1)
#define O2345_ECUDA (-2)
#define O2345_EUNSUPPORTED (-3)

#define O2345_ABI_VERSION 3   /* 2: o2345_epilogue, precision arguments of sdf_query / render_blend, GroupNorm as affine
                                 3: split-K inside the GEMM kernel (cluster per tile, private planes in the workspace), o2345_last_trap, o2345_debug_gemm_force */

typedef void* o2345_stream_t;

int o2345_abi_version(void);
/* Copies the last error message of the calling thread into buf (NUL terminated). */
int o2345_last_error(char* buf, size_t n);
/* Fills (major, minor, sm_count) of the current device; fails if it is not sm_100. */
int o2345_device_info(int* major, int* minor, int* sms);

/* ------------------------------------------------------------------------------------------
 * B8 / B9 / B10: SDF query = quirky trilinear latent fetch + positional embedding +
 * weight-normed 39->128->128->128 MLP (+ analytic d sdf / d x).
 * Replaces SparseSdfNetwork.sdf            reconstruction/models/sparse_sdf_network.py:402-420
 *          ops.grid_sampler.grid_sample_3d reconstruction/ops/grid_sampler.py:64-216
 *          LatentSDFLayer.forward          reconstruction/models/sparse_sdf_network.py:111-136
 *          Embedding.forward               reconstruction/models/embedder.py:81-101
 *          SparseSdfNetwork.gradient       reconstruction/models/sparse_sdf_network.py:476-499
 *          extract_fields (lattice mode)   reconstruction/models/sparse_neus_renderer.py:882-905
 * ------------------------------------------------------------------------------------------ */

/* Number of floats of the packed MLP weights (see o2345_sdf_pack_weights). */
#define O2345_SDF_PE 39
#define O2345_SDF_HID 128
#define O2345_SDF_LAT 16
#define O2345_SDF_IN1 144
/* layout (floats): W0t[39][128] b0[128] W1t[144][128] b1[128] W2t[144][128] b2[128]
 *                  W1[128][144] W0[128][48]   (un-transposed copies for the backward pass)  */
#define O2345_SDF_PACK_FLOATS (39 * 128 + 128 + 2 * (144 * 128 + 128) + 128 * 144 + 128 * 48)

/* w0 [128,39], w1 [128,144], w2 [128,144] are the EFFECTIVE (weight-normed) matrices,
 * row-major as nn.Linear stores them; b* the biases.  Writes the packed blob. */
int o2345_sdf_pack_weights(const float* w0, const float* b0, const float* w1, const float* b1,
                           const float* w2, const float* b2, float* pack, o2345_stream_t stream);

/* Where the query points come from. */
#define O2345_PTS_EXPLICIT 0 /* pts [n,3]                                                   */
#define O2345_PTS_LATTICE 1  /* point i = (lin[i/(R*R)], lin[(i/R)%R], lin[i%R]), n = R^3     */
#define O2345_PTS_RAYS 2     /* point i = o[r] + d[r] * z[r*z_stride + s], r = i / S, s = i % S */

typedef struct o2345_points {
  int mode;
  const float* pts;    /* EXPLICIT: [n,3]                                  */
  const float* lin;    /* LATTICE: [R] coordinates                          */
  int R;               /* LATTICE                                           */
  const float* rays_o; /* RAYS: [n_rays,3]                                  */
  const float* rays_d; /* RAYS: [n_rays,3]                                  */
  const float* z;      /* RAYS: depth of sample s on ray r                  */
  int S;               /* RAYS: samples per ray                             */
  int z_stride;        /* RAYS: row stride of z in floats                   */
} o2345_points;

/* vol_cl: conditional volume, channel-last [D,D,D,16] (voxel (x,y,z) -> ((x*D+y)*D+z)*16).
 * active: optional uint8 [n]; points with active[i]==0 are not evaluated and receive
 *         sdf = inactive_sdf, feat = 0, latent = 0, grad = 0 (reference
 *         sparse_neus_renderer.py:135-139, 229-241).
 * Outputs (any may be NULL): sdf [n], feat [n,127], latent [n,16], grad [n,3].
 * If negate != 0 the sdf output is written as -sdf (extract_fields' u = -sdf). */
#define O2345_SDF_FP32 0      /* fp32 FMA GEMMs                                                                        */
#define O2345_SDF_TC_SPLIT 1  /* forward GEMMs on tensor cores, every operand split into fp16 hi + lo (three MMAs per
                                 product, fp32 accumulate): agrees with the fp32 kernel to ~1e-6                         */
int o2345_sdf_query(const o2345_points* src, int64_t n, const float* vol_cl, int D, const float* wpack,
                    const uint8_t* active, float inactive_sdf, int negate, int precision, float* sdf, float* feat,
                    float* latent, float* grad, o2345_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * B3 / B4 / B5 / B7: cost-volume build.
 * Replaces generate_grid + back_project_sparse_type (called twice) + aggregate_multiview_features
 *          + sparse_to_dense_volume:  reconstruction/ops/generate_grids.py:4-19,
 *          reconstruction/ops/back_project.py:5-86,
 *          reconstruction/models/sparse_sdf_network.py:221-284,321-346.
 * ------------------------------------------------------------------------------------------ */

/* proj [V,4,4] = K @ w2c per view (row-major), origin [3] world position of voxel (0,0,0).
 * mask_bits[D^3]: bit v set iff voxel is inside view v's frustum (|gx|<=1, |gy|<=1, z>0);
 * keep[D^3] = popcount(mask) > min_views.  V <= 32. */
int o2345_frustum_mask(const float* proj, int V, const float* origin, float voxel_size, int D, int sizeH,
                       int sizeW, int min_views, uint32_t* mask_bits, uint8_t* keep, o2345_stream_t stream);

/* Ordered stream compaction: rows[k] = i of the k-th non-zero flag (ascending i), index[i] = k or -1
 * (index may be NULL), *count = number kept.  scratch: o2345_compact_scratch_ints(n) int32. */
int64_t o2345_compact_scratch_ints(int64_t n);
int o2345_compact(const uint8_t* flags, int64_t n, int32_t* rows, int32_t* index, int32_t* count,
                  int32_t* scratch, o2345_stream_t stream);

/* feats_nhwc [V,h,w,16] compressed feature maps (channel-last).  For every kept voxel (rows,
 * *count, at most max_rows) writes cost[row] = [var(16), mean(16)] over the V views; features are
 * NOT masked, counts come from mask_bits (reference sparse_sdf_network.py:234-245). */
int o2345_costvol_gather(const float* feats_nhwc, int V, int h, int w, int sizeH, int sizeW, const float* proj,
                         const float* origin, float voxel_size, int D, const int32_t* rows, const int32_t* count,
                         int64_t max_rows, const uint32_t* mask_bits, float* cost, o2345_stream_t stream);

/* Scatter rows [n,16] into vol_cl [D^3,16] (channel-last), optionally vol_cf [16,D^3] (the
 * reference layout [1,16,X,Y,Z]) and occ [D^3] (1.0 where a row exists).  Outputs are zero-filled first. */
int o2345_dense_scatter(const float* feat, const int32_t* rows, const int32_t* count, int64_t max_rows, int D,
                        float* vol_cl, float* vol_cf, float* occ, o2345_stream_t stream);

/* Nearest occupancy lookup, ATen grid_sample(mode='nearest', align_corners=False) semantics
 * (reference sparse_neus_renderer.py:153-169).  out[i] = 1 iff occ at the nearest voxel > 0. */
int o2345_occ_nearest(const o2345_points* src, int64_t n, const float* occ, int D, uint8_t* out,
                      o2345_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * B6: sparse 3-D convolution stack (torchsparse v1.4.0 semantics) + BatchNorm(batch stats) + ReLU.
 * Replaces spnn.Conv3d / spnn.BatchNorm / spnn.ReLU as used by SparseCostRegNet:
 *          reconstruction/tsparse/modules.py:94-124,259-304.
 * A level is described by a dense lattice index[E^3] (row id or -1), its row list rows[n] (cell
 * ids, ascending) and a device-side count.  Level l+1 has extent E/2+1.
 * ------------------------------------------------------------------------------------------ */

/* Flags the cells of the next coarser level (k=3, stride 2 down-sampling rule). cmin_scratch: int32[3]. */
int o2345_sp_coarsen(const int32_t* fine_index, int Ef, const int32_t* fine_rows, const int32_t* fine_count,
                     int64_t max_fine, int Ec, uint8_t* coarse_flags, int32_t* cmin_scratch,
                     o2345_stream_t stream);

/* mode 0: same level; 1: stride-2 down (in = fine, out = coarse); 2: transposed stride-2 (in = coarse,
 * out = fine).  kernel [27,Cin,Cout] (x-fastest offsets).  Writes raw outputs [rows,Cout] and the
 * per-channel sum / sum of squares into stats[2*Cout] (float64, zeroed by the call). */
int o2345_sp_conv(const float* in_feats, const int32_t* in_index, int Ein, const int32_t* out_rows,
                  const int32_t* out_count, int64_t max_out, int Eout, int mode, const float* kernel, int Cin,
                  int Cout, float* out_raw, double* stats, o2345_stream_t stream);

/* out = relu(batchnorm(x; batch stats, gamma, beta, eps)) (+ skip).  out may alias x. */
int o2345_sp_bn_relu(const float* x, const int32_t* count, int64_t max_rows, int C, const double* stats,
                     const float* gamma, const float* beta, float eps, const float* skip, float* out,
                     o2345_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * B10: marching cubes on the dense u = -sdf grid.
 * Replaces mcubes.marching_cubes(u, 0): reconstruction/models/sparse_neus_renderer.py:932.
 * Case tables come from the host (o2345/mc_tables.py): tri_table int8[256,16], n_tri uint8[256],
 * edge_owner int8[12,4] = (dx,dy,dz,axis) of the lattice edge that carries cell edge e.
 * ------------------------------------------------------------------------------------------ */
int o2345_mc_classify(const float* u, int R, float iso, uint8_t* cases, uint8_t* cell_flags, uint8_t* edge_flags,
                      o2345_stream_t stream);
int o2345_mc_vertices(const float* u, int R, float iso, const int32_t* edges, const int32_t* count,
                      int64_t max_verts, double* verts, o2345_stream_t stream);
int64_t o2345_scan_scratch_ints(int64_t n);
int o2345_mc_tri_offsets(const uint8_t* cases, const int32_t* cells, const int32_t* count, int64_t max_cells,
                         const uint8_t* n_tri_table, int32_t* offsets, int32_t* total, int32_t* scratch,
                         o2345_stream_t stream);
int o2345_mc_triangles(const uint8_t* cases, int R, const int32_t* cells, const int32_t* count, int64_t max_cells,
                       const int32_t* tri_offsets, const int8_t* tri_table, const uint8_t* n_tri_table,
                       const int8_t* edge_owner, const int32_t* vert_index, int32_t* tris, o2345_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * B1 / B2: FeatureNet + compress layer primitives.
 * Replaces nn.Conv2d + InPlaceABN + F.interpolate in reconstruction/models/featurenet.py:12-91,
 *          reconstruction/models/trainer_generic.py:1104-1125, sparse_sdf_network.py:171-173.
 * ------------------------------------------------------------------------------------------ */
typedef struct o2345_view4 {
  float* ptr;               /* element (n,c,h,w) lives at ptr[n*sn + (c+c0)*sc + h*sh + w*sw] */
  int64_t sn, sc, sh, sw;
  int c0;
} o2345_view4;

/* in [N,Cin,H,W] NCHW, weight [Cout,Cin,K,K], bias NULL or [Cout]; out NCHW raw.  If stats != NULL the
 * per-channel sum / sum of squares of the output are accumulated into stats[2*Cout] (zeroed first). */
int o2345_conv2d(const float* in, int N, int Cin, int H, int W, const float* weight, const float* bias, int Cout,
                 int K, int stride, int pad, float* out, double* stats, o2345_stream_t stream);
/* InPlaceABN forward with batch statistics: (x-mean)/sqrt(var+eps)*(|gamma|+eps)+beta, leaky-ReLU(slope). */
int o2345_abn_apply(const float* x, int N, int C, int H, int W, const double* stats, const float* gamma,
                    const float* beta, float eps, float slope, const o2345_view4* out, o2345_stream_t stream);
/* Bilinear up-sampling by an integer factor, align_corners=True, optional add [N,C,H*f,W*f]. */
int o2345_upsample_bilinear(const float* x, int N, int C, int H, int W, int factor, const float* add,
                            const o2345_view4* out, o2345_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * B11 - B14: volume rendering.
 * Replaces SparseNeuSRenderer.up_sample / cat_z_vals / render_core / render
 *              reconstruction/models/sparse_neus_renderer.py:73-151,171-455,457-635
 *          sample_pdf, sample_ptsFeatures_from_feature{Volume,Maps}
 *              reconstruction/models/render_utils.py:8-120
 *          Projector.compute / compute_view_independent / compute_angle*
 *              reconstruction/models/projector.py:15-62,96-425
 *          GeneralRenderingNetwork.forward
 *              reconstruction/models/rendering_network.py:75-129
 * ------------------------------------------------------------------------------------------ */

/* One importance round: new_z [R,n_new] drawn by deterministic inverse-CDF sampling (u [n_new] =
 * linspace(0.5/n, 1-0.5/n, n)) from the NeuS section weights of the current samples z/sdf [R,S]. */
int o2345_ray_upsample(const float* rays_o, const float* rays_d, int64_t R, const float* z, const float* sdf,
                       int S, float inv_s, const float* occ, int D, const float* u, int n_new, float* new_z,
                       o2345_stream_t stream);
/* Merge the sorted lists (z,sdf) [R,S] and (new_z,new_sdf) [R,n_new] into out_* [R,S+n_new]. */
int o2345_ray_merge(const float* z, const float* sdf, int S, const float* new_z, const float* new_sdf, int n_new,
                    int64_t R, float* out_z, float* out_sdf, o2345_stream_t stream);
/* mid_z = z + dists/2, dists = forward differences (last = sample_dist), active = nearest occupancy. */
int o2345_ray_midpoints(const float* rays_o, const float* rays_d, int64_t R, const float* z, int S,
                        float sample_dist, const float* occ, int D, float* mid_z, float* dists, uint8_t* active,
                        o2345_stream_t stream);

#define O2345_MAP_CH 60 /* channel-last source maps: rgb(3) + pyramid features(56) + 1 pad */
#define O2345_RNET_PACK_FLOATS 19664

typedef struct o2345_views {
  int V, H, W;          /* source views and map size                                        */
  const float* maps;    /* [V,H,W,60] channel-last: [0..2] colour, [3..58] features, [59] 0  */
  const float* proj;    /* [V,3,4] = intrinsics @ w2c[:3,:4]                                 */
  const float* centers; /* [V,3] camera centres (c2w translation)                           */
  float sizeW, sizeH;   /* img_wh used to normalise pixel coordinates                       */
} o2345_views;

/* Per sample point: geometry feature, per-view colour+feature fetch, ray-difference, view-blending
 * MLP -> rgb [n,3]; nvalid [n] = number of views whose mask is set (may be NULL).  dir_mode 0: target
 * direction = normalised (query_center - p) (Projector.compute); 1: dirs [n,3] given
 * (compute_view_independent, surface normals).  rnet_pack: O2345_RNET_PACK_FLOATS floats, every
 * matrix stored [in][out] in the order documented in csrc/render.cu. */
#define O2345_BLEND_FP32 0     /* fp32 FMA mat-vecs in the reference's operation order (tight oracle parity)            */
#define O2345_BLEND_TC_FP16 1  /* per-(sample, view) MLPs as mma.sync products: fp16 operands, fp32 accumulate / statistics */
#define O2345_BLEND_TC5 2      /* the same MLPs as tcgen05.mma M = 128 tiles (4 samples x 32 views, a lane is a view), TMEM accumulators */
int o2345_render_blend(const o2345_points* src, int64_t n, const uint8_t* active, const float* vol_cl,
                       const float* occ, int D, const o2345_views* views, int dir_mode, const float* query_center,
                       const float* dirs, const float* rnet_pack, int precision, float* rgb, int32_t* nvalid,
                       o2345_stream_t stream);

/* NeuS alpha from (sdf, grad), transmittance, colour/depth compositing.  Outputs: color [R,3], depth [R],
 * optional weights [R,S], cdf [R,S], alpha [R,S], weights_sum [R], color_mask [R] (uint8). */
int o2345_ray_composite(const float* rays_d, int64_t R, int S, const float* mid_z, const float* dists,
                        const float* sdf, const float* grad, const float* color, const uint8_t* active,
                        const int32_t* nvalid, float inv_s, float alpha_inter_ratio, int has_background,
                        float background, float* color_out, float* depth_out, float* weights_out, float* cdf_out,
                        float* alpha_out, float* weights_sum_out, uint8_t* color_mask_out, o2345_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Path A (rows A2-A4, A6): fp16 tensor-core GEMM, tcgen05.mma + TMEM accumulators + TMA operands.
 * Replaces the cuBLAS / cuDNN calls behind nn.Linear, 1x1 and (im2col'd) 3x3 nn.Conv2d and the
 * attention einsums of the Zero123 UNet and VAE:
 *          ldm/modules/diffusionmodules/openaimodel.py:745-777, ldm/modules/attention.py:170-193,
 *          ldm/modules/diffusionmodules/model.py:535-568.
 * C[b] = act(alpha * A[b] . B[b]^T + bias + rowbias) + residual[b];  A [M,K] (row stride lda), B [N,K] (row
 * stride ldb), fp16, K contiguous; C and residual [M,N] (row stride ldc), C fp16 or fp32.  nh = 0: plain GEMM;
 * nh > 0: nh * nb independent products, operand z = b * nh + h lives at ptr + h * stride_*_h + b * stride_*_b
 * (e.g. heads inside a [B, N, H*d] tensor).  lda, ldb and the A/B batch strides must be multiples of 8 elements
 * (TMA: 16-byte strides).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const float* bias;     /* [N] fp32 or NULL */
  const void* residual;  /* fp16 [M, ldc] or NULL, added after the activation (ResBlock / attention skip) */
  const void* rowbias;   /* fp16 or NULL: element [(row / rows_per_group) * rowbias_ld + col] is added before the
                            activation (the ResBlock's per-image emb_layers output, openaimodel.py:266-273) */
  int64_t rowbias_ld;
  int rows_per_group;
  int act;               /* 0 none, 1 SiLU, 2 GELU(erf), 4 QuickGELU x*sigmoid(1.702x) (CLIP), 3 GEGLU (attention.py:37-44): the N columns come in chunks of
                            32 = 16 values followed by their 16 gates, C gets N/2 columns value * gelu(gate) */
  float alpha;           /* scale on the accumulator */
  int out_f32;           /* C is fp32 instead of fp16 */
  float* colstats;       /* optional fp32 [M / stats_rows_per_group, 2, N], zero on entry: the kernel ADDS, per row group (image)
                            and column, the sum and the sum of squares of the fp16 values it writes -- the statistics the next
                            GroupNorm needs (openaimodel.py:256-276), so that no kernel has to re-read the tensor for them.
                            fp16 output, act 0, N and ldc multiples of 8; stats_rows_per_group 64 or a multiple of 128 */
  int stats_rows_per_group;
} o2345_epilogue;

int o2345_gemm_f16(const void* A, const void* B, void* C, int M, int N, int K, int64_t lda, int64_t ldb,
                   int64_t ldc, int nh, int nb, int64_t stride_a_h, int64_t stride_a_b, int64_t stride_b_h,
                   int64_t stride_b_b, int64_t stride_c_h, int64_t stride_c_b, const o2345_epilogue* ep /* NULL: plain */,
                   float* splitk_ws, int64_t ws_floats, o2345_stream_t stream);
/* splitk_ws (optional, may be NULL): fp32 scratch of ws_floats elements, no initialisation needed.  When the output tiles
 * alone cannot fill the GPU the K range is split over up to min(8, ws_floats / (M*N)) CTA pairs per tile that run as one
 * thread-block cluster: each stores its partial tile in its own [M, N] plane of the scratch, a cluster barrier publishes the
 * planes, and every split sums them and applies the epilogue to its share of the tile.  One workspace serves one stream at
 * a time. */

/* Every mbarrier wait inside the GEMM kernel is bounded (4 s).  If one expires the kernel records which barrier of which CTA
 * of which problem stalled in a host-mapped buffer and traps (the CUDA context then reports a launch failure at the next
 * synchronisation).  o2345_last_trap copies a description of that record into buf and returns 1, or returns 0 if no wait
 * has ever expired in this process.  Safe to call after the context has failed. */
int o2345_last_trap(char* buf, size_t n);

/* Tuning hook (tools/gemm_sweep.py; not part of the data path): force the tile configuration of the following non-batched
 * GEMM / conv calls: ctas in {1, 2}, bn in {64, 128, 160, 256}, splits >= 1; 0 keeps the heuristic's choice of that field. */
void o2345_debug_gemm_force(int ctas, int bn, int splits);
/* Tuning hook: the persistent variant of the kernel (one CTA pair per SM pair walking many tiles, epilogue of tile i under
 * the main loop of tile i + 1).  mode 0: heuristic (at least min_tiles pair tiles; min_tiles 0 = default), 1: wherever it is
 * available (pair tiles of 128+ columns, staged fp16 epilogue, no split-K), 2: never. */
void o2345_debug_gemm_persist(int mode, int min_tiles);
/* Tuning hook: the seven constants of the tile-configuration cost model (per-SM ingest B/clk, two-CTA bonus, fabric B/clk,
 * fixed us, epilogue us per 160 columns, split-K us, split-K us per split and 128 columns); see gemm_tc.cu predict_us. */
void o2345_debug_gemm_model(const float* seven);

/* Diagnostic hook (not part of the data path): when device_buf16 != NULL, CTA (0,0,0) of every following CTA-pair GEMM
 * launch stores clock64() stamps of its phases into device_buf16[0..8] (entry, prologue done, first TMA issued, last TMA
 * issued, first operands landed, last MMA issued, accumulator ready, epilogue done, exit).  NULL switches it off. */
void o2345_debug_gemm_trace(long long* device_buf16);

/* Implicit-GEMM 3x3 convolution, stride 1, zero padding 1 (nn.Conv2d(C, N, 3, padding=1) of the UNet ResBlocks and the
 * VAE ResnetBlocks): x channel-last [B,H,W,C] fp16, weight [N, 9*C] fp16 in (ky,kx,c) order, out [B*H*W, N] (row stride
 * ldc).  No im2col buffer exists: the nine shifted windows are fetched by 4-D TMA boxes whose out-of-bounds zero fill is
 * the padding.  W must divide 128 or be a multiple of 128; C a multiple of 8.  Epilogue as o2345_gemm_f16. */
int o2345_conv3x3_f16(const void* x, int B, int H, int W, int C, const void* weight, int N, void* out, int64_t ldc,
                      const o2345_epilogue* ep, float* splitk_ws, int64_t ws_floats, o2345_stream_t stream);

/* Nearest-neighbour 2x up-sampling followed by a 3x3 convolution (zero padding 1) -- the Upsample layers of the UNet and the
 * VAE decoder (reference ldm/modules/diffusionmodules/openaimodel.py:118-129 `Upsample.forward`, model.py:43-53) -- WITHOUT
 * materialising the up-sampled map or a patch matrix: every output pixel (2y+a, 2x+b) sees the 3x3 kernel collapse onto a
 * 2x2 window of the low-resolution input, so the layer is four 2x2 implicit convolutions, one per phase (a, b).
 * x [B, H, W, C] channel-last fp16 (the LOW-resolution map, same tiling rule as o2345_conv3x3_f16); weight4 [4][N][4*C] fp16:
 * phase 2a+b, taps in (ty, tx, c) order with the collapsed kernel rows / columns summed (a = 0: {k0, k1+k2}, a = 1: {k0+k1, k2});
 * out [B*2H*2W, ldc] fp16.  Epilogue: bias / activation only. */
int o2345_conv_up2x_f16(const void* x, int B, int H, int W, int C, const void* weight4, int N, void* out, int64_t ldc,
                        const o2345_epilogue* ep, float* splitk_ws, int64_t ws_floats, o2345_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Path A glue (rows A1, A3, A4, A6): channel-last fp16 activations [B, H*W, C]; fp32 statistics.
 * GroupNorm32 / SiLU / conv patch gather: ldm/modules/diffusionmodules/openaimodel.py:92-161,256-276,
 *   ldm/modules/diffusionmodules/util.py:214-216; LayerNorm / softmax / GEGLU: ldm/modules/attention.py:37-64,
 *   170-193,214-218; timestep embedding: util.py:151-171; CFG + DDIM update: ldm/models/diffusion/ddim.py:196-243.
 * ------------------------------------------------------------------------------------------ */
/* GroupNorm statistics folded into a per-(image, channel) affine: GroupNorm(x)[b, p, c] = x * scale[b, c] + shift[b, c]
 * (scale = rstd * gamma, shift = beta - mean * rstd * gamma; gamma / beta may be NULL).  x [B, HW, C] fp16, C a multiple
 * of 8.  scratch: o2345_groupnorm_scratch_floats(B, G) fp32 words, ALL ZERO on entry and left zeroed. */
int64_t o2345_groupnorm_scratch_floats(int B, int G);
int o2345_groupnorm_stats(const void* x, int B, int HW, int C, int G, float eps, const float* gamma, const float* beta,
                          float* scratch, float* scale, float* shift, o2345_stream_t stream);
/* GroupNorm(x) (+SiLU if act) of a channel-last activation x [B, HW, C] -> out [B, HW, C], statistics and apply in ONE kernel:
 * each image is handled by a thread-block cluster whose CTAs exchange their partial sums through distributed shared memory
 * (no scratch, no global atomics).  Same result as o2345_groupnorm_stats + o2345_norm_act_im2col(ksize 1). */
int o2345_groupnorm_apply(const void* x, int B, int HW, int C, int G, float eps, const float* gamma, const float* beta, int act,
                          void* out, o2345_stream_t stream);
/* Tuning hook (tools/gn_bench.py): CTAs per image (cluster size, a power of two <= 16) of o2345_groupnorm_apply; 0 = the launcher's rule. */
void o2345_debug_groupnorm_cluster(int cl);
/* out [B*Ho*Wo, k*k*C] (column order ky,kx,c) = patches of f(x), f = x * scale + shift (+SiLU if act) when scale != NULL.
 * upsample != 0: nearest x2 replication of x before the convolution.  Zero padding k/2 on the high side and
 * pad_lo on the low side (pad_lo < 0: k/2; pad_lo = 0 reproduces the VAE encoder's F.pad(x, (0,1,0,1))). */
int o2345_norm_act_im2col(const void* x, int B, int H, int W, int C, int ksize, int stride, int upsample, int pad_lo,
                          const float* scale, const float* shift, int act, void* out, o2345_stream_t stream);
/* The same gather with GroupNorm(x) (+SiLU if act) computed from RAW statistics: stats_a [B, 2, Ca] (sum, then sum of squares,
 * per image and channel, over the H*W pixels of x) for channels [0, Ca) and stats_b [B, 2, C - Ca] for the rest (NULL when
 * Ca == C) -- the tables the producing GEMMs accumulated through o2345_epilogue.colstats (two tables: x is a channel
 * concat).  G groups, eps, gamma / beta [C] (may be NULL).  Replaces o2345_groupnorm_stats + o2345_norm_act_im2col. */
int o2345_norm_act_im2col_stats(const void* x, int B, int H, int W, int C, int ksize, int stride, int upsample, int pad_lo,
                                const float* stats_a, int Ca, const float* stats_b, int G, float eps, const float* gamma,
                                const float* beta, int act, void* out, o2345_stream_t stream);
int o2345_layernorm_rows(const void* x, int64_t M, int C, float eps, const float* gamma, const float* beta, void* y,
                         o2345_stream_t stream);
int o2345_softmax_rows(const void* s, int64_t rows, int n, void* p, o2345_stream_t stream);
/* Fused multi-head self-attention (ldm/modules/attention.py:170-193): out[b, n, h*d + j] = softmax(q k^T * scale) v.
 * q, k, v: fp16 [B*N, >= H*d] views with a common row stride ld (e.g. column blocks of a fused qkv projection);
 * d in {40, 64, 80, 160} (64: CLIP ViT-L/14); scores stay on chip (mma.sync m16n8k16, fp32 online softmax). */
int o2345_attention_f16(const void* q, const void* k, const void* v, int B, int N, int H, int d, int ld, void* out,
                        int ldo, float scale, o2345_stream_t stream);
/* y[M,I] = x[:, :I] * gelu(x[:, I:2I]) */
int o2345_geglu(const void* x, int64_t M, int I, void* y, o2345_stream_t stream);
int o2345_silu(const void* x, int64_t n, void* y, o2345_stream_t stream);
int o2345_transpose_tokens(const void* x, int B, int N, int C, void* y, o2345_stream_t stream);
int o2345_timestep_embedding(const float* t, int B, int dim, void* out, o2345_stream_t stream);
/* y[b, p, c] += e[b * lde + c] */
int o2345_add_channel_bias(void* y, const void* e, int B, int HW, int C, int lde, o2345_stream_t stream);
int o2345_copy_channels(const void* src, int64_t M, int C, void* dst, int ldd, int off, o2345_stream_t stream);
int o2345_nchw_f32_to_cl_f16(const float* x, int B, int C, int HW, void* y, int ldy, int off, o2345_stream_t stream);
int o2345_cl_f16_to_nchw_f32(const void* x, int B, int C, int HW, int ldx, float* y, o2345_stream_t stream);
/* eps = [unconditional | conditional] halves of n elements each; writes x_prev and (optionally) pred_x0. */
int o2345_cfg_ddim_update(const float* x, const float* eps, const float* noise, int64_t n, float scale, float a_t,
                          float a_prev, float sigma_t, float sqrt_one_minus_at, float* x_prev, float* pred_x0,
                          o2345_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Row A8: front end of the CLIP ViT-L/14 image tower (FrozenCLIPImageEmbedder.preprocess + the patch embedding's patch
 * gather, ldm/modules/encoders/modules.py:362-370).  mean3 / std3 are HOST pointers to three floats.
 * out [B * (res/patch)^2, kp] fp16: row (b, py, px), column (c, ky, kx) = bicubic(align_corners) resize of x [B,3,H,W]
 * (fp32, [-1,1]) to res x res, mapped to [0,1] and normalised; columns 3 patch^2 .. kp-1 zero.
 * ------------------------------------------------------------------------------------------ */
int o2345_clip_patches(const float* x, int B, int H, int W, int res, int patch, const float* mean3, const float* std3, int kp,
                       void* out, o2345_stream_t stream);
/* tok [B*N, d] fp16: row 0 of every image := class_embedding + pos[0]; rows n >= 1 += pos[n] (cls, pos fp32 on the device) */
int o2345_clip_add_positions(void* tok, const float* cls, const float* pos, int B, int N, int d, o2345_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* O2345_H_ */
