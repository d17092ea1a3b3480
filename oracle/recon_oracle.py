"""CPU oracle for the SparseNeuS-style reconstruction hot path.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package; the product
(one-2-3-45_b200/) never does and fails loudly when its CUDA library is missing.

It is a standalone fp32 torch-CPU restatement of the reference's algorithm (the reference
is pure Python and cannot travel to the GPU box).  Every function cites the reference
file:line it follows.  Pinning status:

* functions restating code that lives under /root/reference are checked against the
  reference's own Python by oracle/pin_against_reference.py (run in the build container,
  results frozen under tests/golden/);
* `torchsparse_conv3d`, `inplace_abn` and `marching_cubes` restate third-party packages
  whose source is NOT in /root/reference (torchsparse v1.4.0, inplace_abn, PyMCubes>=0.1.4;
  reference README.md:102-105, requirements.txt:52) -- PARITY UNPINNED for those three.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# B3/B4/B5  lattice, back-projection, variance/mean cost
# --------------------------------------------------------------------------------------


def lattice_coords(dim):
    """[D^3, 3] float32 voxel indices, x-major (reference ops/generate_grids.py:4-19 +
    sparse_sdf_network.py:321-326: meshgrid 'ij' flattened as x*D*D + y*D + z)."""
    r = torch.arange(dim, dtype=torch.float32)
    g = torch.stack(torch.meshgrid(r, r, r, indexing="ij"), -1)
    return g.reshape(-1, 3)


def project_voxels(coords, origin, voxel_size, proj, H, W):
    """Voxel -> normalised image grid for every view (reference ops/back_project.py:44-61).

    coords [N,3] float voxel indices, proj [V,4,4] (K @ w2c).  Returns gx, gy, z [V,N]
    and the int32 visibility mask [N,V].  The 4-term dot products are evaluated left to
    right with separately rounded multiplies and adds; the CUDA kernel uses the same order
    (__fmul_rn/__fadd_rn) so masks can be compared bit for bit.
    """
    vs = torch.tensor(voxel_size, dtype=torch.float32)
    w = coords * vs + origin.float()[None]
    x, y, z = w[:, 0][None], w[:, 1][None], w[:, 2][None]
    P = proj.float()

    def row(r):
        return ((P[:, r, 0:1] * x + P[:, r, 1:2] * y) + P[:, r, 2:3] * z) + P[:, r, 3:4]

    ix, iy, iz = row(0), row(1), row(2)
    iz = torch.where(iz >= 0, iz.clamp(min=1e-6), iz)
    u, v = ix / iz, iy / iz
    gx = 2 * u / float(W - 1) - 1
    gy = 2 * v / float(H - 1) - 1
    mask = (gx.abs() <= 1) & (gy.abs() <= 1) & (iz > 0)
    return gx, gy, iz, mask.t().contiguous().to(torch.int32)


def frustum_keep(mask, min_views=1):
    """Voxels seen by more than `min_views` views (reference sparse_sdf_network.py:303,333)."""
    return mask.sum(-1) > min_views


def backproject_features(coords, origin, voxel_size, feats, proj, H, W):
    """Per-view bilinear fetch, NOT masked (reference ops/back_project.py:70-78).
    feats [V,C,h,w] -> [N,V,C], mask [N,V]."""
    gx, gy, _, mask = project_voxels(coords, origin, voxel_size, proj, H, W)
    grid = torch.stack([gx, gy], -1)[:, None]            # [V,1,N,2]
    f = F.grid_sample(feats, grid, padding_mode="zeros", align_corners=True)  # [V,C,1,N]
    return f[:, :, 0].permute(2, 0, 1).contiguous(), mask


def variance_mean(feats, mask):
    """[N,V,C],[N,V] -> [N,2C] = cat(var, mean) (reference sparse_sdf_network.py:221-250)."""
    cnt = mask.sum(1).float()
    inv = 1.0 / (cnt + 1e-5)
    s = feats.sum(1)
    sq = (feats ** 2).sum(1)
    mean = s * inv[:, None]
    return torch.cat([sq * inv[:, None] - mean ** 2, mean], 1)


# --------------------------------------------------------------------------------------
# third-party restatements (PARITY UNPINNED): inplace_abn, torchsparse v1.4.0
# --------------------------------------------------------------------------------------


def inplace_abn(x, gamma, beta, eps=1e-5, slope=0.01):
    """InPlaceABN forward in training mode: batch statistics over (N,H,W), biased variance,
    scale |gamma|+eps, leaky-ReLU (call sites reference featurenet.py:19-22,
    sparse_sdf_network.py:171-173).  The reference never calls .eval()."""
    dims = [0] + list(range(2, x.dim()))
    mean = x.mean(dims, keepdim=True)
    var = x.var(dims, unbiased=False, keepdim=True)
    shape = [1, -1] + [1] * (x.dim() - 2)
    y = (x - mean) / torch.sqrt(var + eps) * (gamma.abs() + eps).view(shape) + beta.view(shape)
    return F.leaky_relu(y, slope)


def batchnorm_rows(x, gamma, beta, eps=1e-5):
    """nn.BatchNorm1d in training mode on [N,C] rows (spnn.BatchNorm, reference tsparse/modules.py:103)."""
    mean = x.mean(0, keepdim=True)
    var = x.var(0, unbiased=False, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps) * gamma[None] + beta[None]


def kernel_offsets(tensor_stride):
    """27 offsets, x fastest (torchsparse v1.4.0 get_kernel_offsets for odd kernel volume)."""
    r = [-1, 0, 1]
    return torch.tensor([[x, y, z] for z in r for y in r for x in r], dtype=torch.int64) * tensor_stride


def _coord_key(c, lo, ext):
    c = c - lo
    return (c[:, 0] * ext + c[:, 1]) * ext + c[:, 2]


def downsample_coords(coords, tensor_stride):
    """Output coordinates of a k=3, stride-2 sparse conv (torchsparse v1.4.0 spdownsample,
    kernel!=stride branch): every input+offset whose components are multiples of 2*ts and
    >= the per-axis minimum of the inputs, de-duplicated and sorted lexicographically."""
    off = kernel_offsets(tensor_stride)
    cmin = coords.min(0, keepdim=True).values
    cand = (coords[:, None, :] + off[None]).reshape(-1, 3)
    ok = ((cand % (2 * tensor_stride)) == 0).all(1) & (cand >= cmin).all(1)
    return torch.unique(cand[ok], dim=0)


def kernel_map(in_coords, out_coords, tensor_stride):
    """idx[j,k] = row of the input voxel at out_coords[j] + offset_k, or -1."""
    lo = torch.minimum(in_coords.min(0).values, out_coords.min(0).values) - tensor_stride
    hi = torch.maximum(in_coords.max(0).values, out_coords.max(0).values) + tensor_stride
    ext = int((hi - lo).max().item()) + 1
    table = torch.full((ext ** 3,), -1, dtype=torch.int64)
    table[_coord_key(in_coords, lo, ext)] = torch.arange(in_coords.shape[0])
    off = kernel_offsets(tensor_stride)
    q = (out_coords[:, None, :] + off[None]).reshape(-1, 3)
    return table[_coord_key(q, lo, ext)].reshape(out_coords.shape[0], 27)


def torchsparse_conv3d(feats, coords, tensor_stride, kernel, stride, transposed, cmaps, kmaps):
    """spnn.Conv3d(k=3) forward (torchsparse v1.4.0 semantics).

    coords int [N,4] = (x,y,z,b) with a single batch.  Returns (feats, coords, stride,
    cmaps, kmaps).  out[j] = sum_k in[idx(coord_j + off_k)] @ kernel[k]; a transposed conv
    reuses the cached map of the matching down-conv with the roles of its two columns
    swapped and the same kernel index.
    """
    xyz = coords[:, :3].long()
    cmaps = dict(cmaps)
    kmaps = dict(kmaps)
    cmaps.setdefault(tensor_stride, coords)
    if not transposed:
        if stride == 1:
            out_xyz, out_stride = xyz, tensor_stride
        else:
            out_stride = tensor_stride * stride
            if out_stride in cmaps:
                out_xyz = cmaps[out_stride][:, :3].long()
            else:
                out_xyz = downsample_coords(xyz, tensor_stride)
        key = (tensor_stride, stride)
        if key not in kmaps:
            kmaps[key] = kernel_map(xyz, out_xyz, tensor_stride)
        idx = kmaps[key]
        out = torch.zeros(out_xyz.shape[0], kernel.shape[2], dtype=feats.dtype)
        for k in range(27):
            j = torch.nonzero(idx[:, k] >= 0).squeeze(1)
            if j.numel():
                out.index_add_(0, j, feats[idx[j, k]] @ kernel[k])
        out_coords = torch.cat([out_xyz.to(coords.dtype), coords[:1, 3:].expand(out_xyz.shape[0], 1)], 1)
        cmaps.setdefault(out_stride, out_coords)
        return out, out_coords, out_stride, cmaps, kmaps
    fine_stride = tensor_stride // stride
    idx = kmaps[(fine_stride, stride)]                 # [n_coarse, 27] -> fine rows
    fine_coords = cmaps[fine_stride]
    out = torch.zeros(fine_coords.shape[0], kernel.shape[2], dtype=feats.dtype)
    for k in range(27):
        j = torch.nonzero(idx[:, k] >= 0).squeeze(1)
        if j.numel():
            out.index_add_(0, idx[j, k], feats[j] @ kernel[k])
    return out, fine_coords, fine_stride, cmaps, kmaps


def cost_reg_net(feats, xyz, sd, prefix="sparse_costreg_net."):
    """SparseCostRegNet forward (reference tsparse/modules.py:287-304) on rows `feats`
    [N,Cin] with integer voxel coordinates xyz [N,3]; conv -> batch-stat BN -> ReLU blocks."""
    coords = torch.cat([xyz.int(), torch.zeros(xyz.shape[0], 1, dtype=torch.int32)], 1)
    state = {"cm": {}, "km": {}}

    def block(name, x, c, s, stride=1, transposed=False):
        k = sd[prefix + name + ".net.0.kernel"]
        y, c2, s2, state["cm"], state["km"] = torchsparse_conv3d(x, c, s, k, stride, transposed,
                                                                   state["cm"], state["km"])
        y = batchnorm_rows(y, sd[prefix + name + ".net.1.weight"], sd[prefix + name + ".net.1.bias"])
        return torch.relu(y), c2, s2

    c0, cc0, s0 = block("conv0", feats, coords, 1)
    x, c, s = block("conv1", c0, cc0, s0, 2)
    c2, cc2, s2 = block("conv2", x, c, s)
    x, c, s = block("conv3", c2, cc2, s2, 2)
    c4, cc4, s4 = block("conv4", x, c, s)
    x, c, s = block("conv5", c4, cc4, s4, 2)
    x, c, s = block("conv6", x, c, s)
    x, c, s = block("conv7", x, c, s, 2, True)
    x = c4 + x
    x, c, s = block("conv9", x, c, s, 2, True)
    x = c2 + x
    x, c, s = block("conv11", x, c, s, 2, True)
    return c0 + x


def sparse_to_dense(xyz, feats, dim):
    """Scatter rows into [1,C,D,D,D] + occupancy [1,1,D,D,D] (reference
    sparse_sdf_network.py:252-284, tsparse/torchsparse_utils.py:125-130)."""
    C = feats.shape[1]
    dense = torch.zeros(dim, dim, dim, C)
    occ = torch.zeros(dim, dim, dim, 1)
    x = xyz.long()
    dense[x[:, 0], x[:, 1], x[:, 2]] = feats
    occ[x[:, 0], x[:, 1], x[:, 2]] = 1.0
    return dense.permute(3, 0, 1, 2).contiguous()[None], occ.permute(3, 0, 1, 2).contiguous()[None]


# --------------------------------------------------------------------------------------
# B1/B2  FeatureNet + compress layer
# --------------------------------------------------------------------------------------


def _cbr(x, sd, name, stride, pad):
    y = F.conv2d(x, sd[name + ".conv.weight"], None, stride, pad)
    return inplace_abn(y, sd[name + ".bn.weight"], sd[name + ".bn.bias"])


def feature_net(imgs, sd):
    """FPN features [feat2, feat1, feat0] (reference featurenet.py:74-91)."""
    c0 = _cbr(_cbr(imgs, sd, "conv0.0", 1, 1), sd, "conv0.1", 1, 1)
    c1 = _cbr(_cbr(_cbr(c0, sd, "conv1.0", 2, 2), sd, "conv1.1", 1, 1), sd, "conv1.2", 1, 1)
    c2 = _cbr(_cbr(_cbr(c1, sd, "conv2.0", 2, 2), sd, "conv2.1", 1, 1), sd, "conv2.2", 1, 1)
    up = lambda t: F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=True)
    f2 = F.conv2d(c2, sd["toplayer.weight"], sd["toplayer.bias"])
    f1 = up(f2) + F.conv2d(c1, sd["lat1.weight"], sd["lat1.bias"])
    f0 = up(f1) + F.conv2d(c0, sd["lat0.weight"], sd["lat0.bias"])
    f1 = F.conv2d(f1, sd["smooth1.weight"], sd["smooth1.bias"], padding=1)
    f0 = F.conv2d(f0, sd["smooth0.weight"], sd["smooth0.bias"], padding=1)
    return [f2, f1, f0]


def pyramid_feature_maps(imgs, sd):
    """[V,56,H,W] fused pyramid (reference trainer_generic.py:1104-1125)."""
    f2, f1, f0 = feature_net(imgs, sd)
    up = lambda t, s: F.interpolate(t, scale_factor=s, mode="bilinear", align_corners=True)
    return torch.cat([up(f2, 4), up(f1, 2), f0], 1)


def compress_features(fmaps, sd):
    """conv3x3 56->16 + InPlaceABN (reference sparse_sdf_network.py:171-173,311-315)."""
    return _cbr(fmaps, sd, "compress_layer", 1, 1)


def conditional_volume(fmaps, origin, proj, sd, dim, voxel_size, H, W):
    """get_conditional_volume at lod 0 (reference sparse_sdf_network.py:286-400).
    Returns dict(dense, occ, keep, rows) with dense [1,16,D,D,D], occ [1,1,D,D,D]."""
    feats = compress_features(fmaps, sd)
    coords = lattice_coords(dim)
    _, _, _, mask = project_voxels(coords, origin, voxel_size, proj, H, W)
    keep = frustum_keep(mask, min(1, proj.shape[0] - 1))
    kept = coords[keep]
    mv, mm = backproject_features(kept, origin, voxel_size, feats, proj, H, W)
    cost = variance_mean(mv, mm)
    rows = cost_reg_net(cost, kept, sd)
    dense, occ = sparse_to_dense(kept, rows, dim)
    return {"dense": dense, "occ": occ, "keep": keep, "rows": rows, "cost": cost, "mask": mask}


# --------------------------------------------------------------------------------------
# B8/B9  SDF query: quirky trilinear + positional embedding + weight-normed MLP
# --------------------------------------------------------------------------------------


def trilinear_latent(volume, pts):
    """Reference-specific trilinear fetch (reference ops/grid_sampler.py:64-216 as called
    from sparse_sdf_network.py:402-410 after the xyz->zyx flip).

    volume [1,C,X,Y,Z] (cube), pts [n,3] in (x,y,z).  Per axis t = (p+1)/2*(D-1); the point
    is in bounds iff 0 < t < D on all axes (strict at 0, loose at the top); weights come
    from the unclamped floor corners, the corner indices are clamped to [0,D-1] afterwards,
    and out-of-bounds points return zeros.  Differentiable w.r.t. pts through the weights.
    """
    _, C, D, _, _ = volume.shape
    t = (pts + 1) / 2 * (D - 1)                      # [n,3] indexes (X,Y,Z)
    inb = ((t > 0) & (t < D)).all(1)
    f = torch.floor(t).detach()
    w1 = t - f                                       # weight of the +1 corner
    w0 = (f + 1) - t
    vol = volume.reshape(C, -1)
    out = torch.zeros(pts.shape[0], C, dtype=volume.dtype)
    for dx in (0, 1):
        for dy in (0, 1):
            for dz in (0, 1):
                ix = (f[:, 0] + dx).clamp(0, D - 1).long()
                iy = (f[:, 1] + dy).clamp(0, D - 1).long()
                iz = (f[:, 2] + dz).clamp(0, D - 1).long()
                w = (w1[:, 0] if dx else w0[:, 0]) * (w1[:, 1] if dy else w0[:, 1]) * (w1[:, 2] if dz else w0[:, 2])
                out = out + vol[:, (ix * D + iy) * D + iz].t() * w[:, None]
    return torch.where(inb[:, None], out, torch.zeros_like(out))


def positional_embedding(x, n_freqs=6):
    """[x, sin(2^k x), cos(2^k x)]_k (reference models/embedder.py:81-101)."""
    out = [x]
    for k in range(n_freqs):
        fr = float(2 ** k)
        out += [torch.sin(fr * x), torch.cos(fr * x)]
    return torch.cat(out, -1)


def effective_weight(sd, prefix):
    """weight_norm(dim=0): W = g * v / ||v||_row (reference sparse_sdf_network.py:100-101)."""
    v, g = sd[prefix + ".weight_v"], sd[prefix + ".weight_g"]
    return v * (g / v.norm(dim=1, keepdim=True))


def sdf_mlp(pts, latent, sd, prefix="sdf_layer."):
    """LatentSDFLayer.forward for n_layers=4 (reference sparse_sdf_network.py:111-136):
    39 -> 128 -softplus100-> (+latent) 144 -> 128 -softplus100-> (+latent) 144 -> 128."""
    x = positional_embedding(pts)
    x = F.softplus(F.linear(x, effective_weight(sd, prefix + "lin0"), sd[prefix + "lin0.bias"]), beta=100)
    x = torch.cat([x, latent], 1)
    x = F.softplus(F.linear(x, effective_weight(sd, prefix + "lin1"), sd[prefix + "lin1.bias"]), beta=100)
    x = torch.cat([x, latent], 1)
    return F.linear(x, effective_weight(sd, prefix + "lin2"), sd[prefix + "lin2.bias"])


def sdf_query(pts, volume, sd):
    """SparseSdfNetwork.sdf (reference sparse_sdf_network.py:402-420) -> sdf [n,1], feat [n,127], latent [n,16]."""
    latent = trilinear_latent(volume, pts)
    y = sdf_mlp(pts, latent, sd)
    return y[:, :1], y[:, 1:], latent


def sdf_gradient(pts, volume, sd):
    """SparseSdfNetwork.gradient (reference sparse_sdf_network.py:476-499): d sdf / d pts, [n,3]."""
    p = pts.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        s, _, _ = sdf_query(p, volume, sd)
        (g,) = torch.autograd.grad(s.sum(), p)
    return g.detach()


# --------------------------------------------------------------------------------------
# B13  hierarchical sampling
# --------------------------------------------------------------------------------------


def nearest_occupancy(pts, occ):
    """F.grid_sample(mode='nearest', align_corners=False) of the occupancy volume after the
    xyz->zyx flip (reference sparse_neus_renderer.py:153-169).  Closed form per axis:
    idx = round_half_even(((p+1)*D - 1)/2), zero when any idx is outside [0, D-1]."""
    g = torch.flip(pts, dims=[-1]).view(1, 1, 1, -1, 3)
    return F.grid_sample(occ, g, mode="nearest", align_corners=False).view(-1)


def inverse_cdf_samples(bins, weights, n):
    """Deterministic inverse-CDF sampling (reference models/render_utils.py:8-51, det=True)."""
    w = weights + 1e-5
    pdf = w / w.sum(-1, keepdim=True)
    cdf = torch.cat([torch.zeros_like(pdf[:, :1]), torch.cumsum(pdf, -1)], -1)
    u = torch.linspace(0.5 / n, 1.0 - 0.5 / n, n).expand(cdf.shape[0], n).contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = (inds - 1).clamp(min=0)
    above = inds.clamp(max=cdf.shape[-1] - 1)
    c0, c1 = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    b0, b1 = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    den = c1 - c0
    den = torch.where(den < 1e-5, torch.ones_like(den), den)
    return b0 + (u - c0) / den * (b1 - b0)


def importance_round(rays_o, rays_d, z, sdf, n_new, inv_s, occ):
    """SparseNeuSRenderer.up_sample (reference sparse_neus_renderer.py:73-115)."""
    R, S = z.shape
    pts = rays_o[:, None] + rays_d[:, None] * z[..., None]
    m = nearest_occupancy(pts.reshape(-1, 3), occ).reshape(R, S)
    m = m[:, :-1] * m[:, 1:]
    ps, ns = sdf[:, :-1], sdf[:, 1:]
    pz, nz = z[:, :-1], z[:, 1:]
    mid = (ps + ns) * 0.5
    dot = (ns - ps) / (nz - pz + 1e-5)
    prev = torch.cat([torch.zeros(R, 1), dot[:, :-1]], -1)
    dot = torch.minimum(prev, dot).clip(-10.0, 0.0) * m
    dist = nz - pz
    pc = torch.sigmoid((mid - dot * dist * 0.5) * inv_s)
    nc = torch.sigmoid((mid + dot * dist * 0.5) * inv_s)
    alpha = m * ((pc - nc + 1e-5) / (pc + 1e-5))
    T = torch.cumprod(torch.cat([torch.ones(R, 1), 1.0 - alpha + 1e-7], -1), -1)[:, :-1]
    return inverse_cdf_samples(z, alpha * T, n_new)


def merge_samples(rays_o, rays_d, z, new_z, sdf, volume, occ, sd):
    """SparseNeuSRenderer.cat_z_vals (reference sparse_neus_renderer.py:117-151): SDF only
    where the occupancy lookup is positive (100 elsewhere), then sort by depth."""
    R, n_new = new_z.shape
    pts = (rays_o[:, None] + rays_d[:, None] * new_z[..., None]).reshape(-1, 3)
    m = nearest_occupancy(pts, occ) > 0
    new_sdf = torch.full((R * n_new, 1), 100.0)
    if m.float().sum() > 1:
        new_sdf[m] = sdf_query(pts[m], volume, sd)[0]
    zc = torch.cat([z, new_z], -1)
    sc = torch.cat([sdf, new_sdf.view(R, n_new)], -1)
    zc, idx = torch.sort(zc, -1)
    return zc, torch.gather(sc, 1, idx)


def hierarchical_z(rays_o, rays_d, near, far, volume, occ, sd, n_samples=64, n_importance=64, z_init=None):
    """Coarse + 4 importance rounds of SparseNeuSRenderer.render (reference
    sparse_neus_renderer.py:484-545), perturb = 0 unless z_init carries jittered depths."""
    R = rays_o.shape[0]
    if z_init is None:
        z = near + (far - near) * torch.linspace(0.0, 1.0, n_samples)[None]
        z = z.expand(R, n_samples).contiguous()
    else:
        z = z_init
    if n_importance > 0:
        pts = (rays_o[:, None] + rays_d[:, None] * z[..., None]).reshape(-1, 3)
        sdf = sdf_query(pts, volume, sd)[0].reshape(R, n_samples)
        for i in range(4):
            nz = importance_round(rays_o, rays_d, z, sdf, n_importance // 4, 64 * 2 ** i, occ)
            z, sdf = merge_samples(rays_o, rays_d, z, nz, sdf, volume, occ, sd)
    return z


# --------------------------------------------------------------------------------------
# B11/B12/B14  projector, view-blending network, compositing
# --------------------------------------------------------------------------------------


def sample_volume_zeros(pts, volume):
    """ATen trilinear (zeros padding, align_corners=True) + strict |p|<1 validity
    (reference models/render_utils.py:54-85).  pts [n,3], volume [C,X,Y,Z] -> [n,C], [n]."""
    valid = (pts.abs() < 1.0).all(-1)
    g = torch.flip(pts, dims=[-1]).view(1, 1, 1, -1, 3)
    f = F.grid_sample(volume[None], g, padding_mode="zeros", align_corners=True)
    return f.view(volume.shape[0], -1).t(), valid


def sample_views(pts, maps, w2cs, intrinsics, W, H):
    """Bilinear fetch from every view (reference render_utils.py:88-120 + ops/back_project.py:89-129).
    pts [n,3], maps [V,C,h,w] -> feats [V,n,C], mask [V,n]."""
    P = torch.matmul(intrinsics, w2cs[:, :3, :])
    pc = P[:, :3, :3] @ pts.t()[None] + P[:, :3, 3:]
    z = pc[:, 2].clamp(min=1e-3)
    gx = 2 * (pc[:, 0] / z) / (W - 1) - 1
    gy = 2 * (pc[:, 1] / z) / (H - 1) - 1
    gx = torch.where((gx > 1) | (gx < -1), torch.full_like(gx, 2.0), gx)
    gy = torch.where((gy > 1) | (gy < -1), torch.full_like(gy, 2.0), gy)
    mask = (gx.abs() < 1.0) & (gy.abs() < 1.0)
    grid = torch.stack([gx, gy], -1)[:, None]
    f = F.grid_sample(maps, grid, padding_mode="zeros", align_corners=True)[:, :, 0]
    return f.permute(0, 2, 1).contiguous(), mask


def ray_difference(pts, target_dir, cam_centers):
    """[V,n,4]: normalised (target_dir - dir_to_camera) and their dot product
    (reference models/projector.py:15-62).  target_dir [n,3] or broadcastable."""
    to_cam = cam_centers[:, None, :] - pts[None]
    to_cam = to_cam / (to_cam.norm(dim=-1, keepdim=True) + 1e-6)
    diff = target_dir[None] - to_cam
    nrm = diff.norm(dim=-1, keepdim=True)
    dot = (target_dir[None] * to_cam).sum(-1, keepdim=True)
    return torch.cat([diff / nrm.clamp(min=1e-6), dot], -1)


def projector_features(pts, volume, occ, fmaps, imgs, w2cs, intrinsics, W, H):
    """Shared part of Projector.compute / compute_view_independent (reference
    models/projector.py:96-228): geometry feature, per-view rgb+features, validity mask."""
    geo, v0 = sample_volume_zeros(pts, volume[0])
    m1, _ = sample_volume_zeros(pts, occ[0])
    gmask = v0 & (m1[:, 0] > 0)
    feat, vmask = sample_views(pts, fmaps, w2cs, intrinsics, W, H)
    rgb, _ = sample_views(pts, imgs, w2cs, intrinsics, W, H)
    return geo, torch.cat([rgb, feat], -1), gmask[None] & vmask


def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def blend_colors(geo, rgb_feat, ray_diff, mask, sd):
    """GeneralRenderingNetwork.forward for flat points (reference models/rendering_network.py:75-129).
    geo [n,16], rgb_feat [V,n,59], ray_diff [V,n,4], mask [V,n] -> rgb [n,3], n_valid_views [n]."""
    rf = rgb_feat.permute(1, 0, 2)
    rd = ray_diff.permute(1, 0, 2)
    m = mask.permute(1, 0)[..., None].float()
    V = rf.shape[1]
    d = F.elu(_lin(sd, "ray_dir_fc.2", F.elu(_lin(sd, "ray_dir_fc.0", rd))))
    rgb_in = rf[..., :3]
    rf = rf + d
    e = torch.exp(sd["s"].abs() * (rd[..., 3:] - 1))
    w = (e - e.min(1, keepdim=True)[0]) * m
    w = w / (w.sum(1, keepdim=True) + 1e-8)
    mean = (rf * w).sum(1, keepdim=True)
    var = (w * (rf - mean) ** 2).sum(1, keepdim=True)
    x = torch.cat([geo[:, None].expand(-1, V, -1), mean.expand(-1, V, -1), var.expand(-1, V, -1), rf], -1)
    x = F.elu(_lin(sd, "base_fc.2", F.elu(_lin(sd, "base_fc.0", x))))
    xv = F.elu(_lin(sd, "vis_fc.2", F.elu(_lin(sd, "vis_fc.0", x * w))))
    res, vis = xv[..., :-1], xv[..., -1:]
    vis = torch.sigmoid(vis) * m
    x = x + res
    vis = torch.sigmoid(_lin(sd, "vis_fc2.2", F.elu(_lin(sd, "vis_fc2.0", x * vis)))) * m
    x = torch.cat([x, vis, rd], -1)
    x = _lin(sd, "rgb_fc.4", F.elu(_lin(sd, "rgb_fc.2", F.elu(_lin(sd, "rgb_fc.0", x)))))
    x = x.masked_fill(m == 0, -1e9)
    bw = F.softmax(x, 1)
    return (rgb_in * bw).sum(1), m.sum(1)[:, 0]


def vertex_colors(verts, volume, occ, fmaps, imgs, w2cs, intrinsics, sd_sdf, sd_ren, W=256, H=256):
    """compute_view_independent + rendering network on mesh vertices (reference
    models/projector.py:231-425, trainer_generic.py:1338-1361).  All views are supporting views."""
    geo, rgb_feat, mask = projector_features(verts, volume, occ, fmaps, imgs, w2cs, intrinsics, W, H)
    g = sdf_gradient(verts, volume, sd_sdf)
    nrm = F.normalize(g, p=2, dim=-1, eps=1e-6)
    centers = torch.inverse(w2cs)[:, :3, 3]
    rd = ray_difference(verts, nrm, centers)
    rgb, _ = blend_colors(geo, rgb_feat, rd, mask, sd_ren)
    return rgb, nrm


def render_fine(rays_o, rays_d, z, sample_dist, volume, occ, fmaps, imgs, w2cs, intrinsics, query_c2w,
                sd_sdf, sd_ren, variance, W=256, H=256, alpha_inter_ratio=1.0, background_rgb=1.0):
    """SparseNeuSRenderer.render_core with general rendering (reference
    sparse_neus_renderer.py:171-455).  Returns dict(color, depth, weights, gradients, sdf,
    inside, mid_z, color_mask)."""
    R, S = z.shape
    dists = torch.cat([z[:, 1:] - z[:, :-1], torch.full((R, 1), float(sample_dist))], -1)
    mid = z + dists * 0.5
    pts = (rays_o[:, None] + rays_d[:, None] * mid[..., None]).reshape(-1, 3)
    dirs = rays_d[:, None].expand(R, S, 3).reshape(-1, 3)
    pm = nearest_occupancy(pts, occ)
    mb = pm > 0
    if mb.float().sum() < 1:
        mb[:100] = True
    sdf = torch.full((R * S, 1), 100.0)
    feat = torch.zeros(R * S, 127)
    grad = torch.zeros(R * S, 3)
    s, f, _ = sdf_query(pts[mb], volume, sd_sdf)
    sdf[mb], feat[mb] = s, f
    grad[mb] = sdf_gradient(pts[mb], volume, sd_sdf)
    geo, rgb_feat, vmask = projector_features(pts, volume, occ, fmaps, imgs, w2cs, intrinsics, W, H)
    centers = torch.inverse(w2cs)[:, :3, 3]
    to_q = query_c2w.view(-1, 4, 4)[0, :3, 3][None] - pts
    to_q = to_q / (to_q.norm(dim=-1, keepdim=True) + 1e-6)
    rd = ray_difference(pts, to_q, centers)
    color_pts, nvalid = blend_colors(geo, rgb_feat, rd, vmask, sd_ren)
    inv_s = torch.exp(variance * 10.0).clip(1e-6, 1e6)
    cosv = (dirs * grad).sum(-1, keepdim=True)
    it = -(F.relu(-cosv * 0.5 + 0.5) * (1.0 - alpha_inter_ratio) + F.relu(-cosv) * alpha_inter_ratio)
    it = it * pm.view(-1, 1)
    d = dists.reshape(-1, 1)
    nxt = sdf + it.clip(-10.0, 10.0) * d * 0.5
    prv = sdf - it.clip(-10.0, 10.0) * d * 0.5
    pc, nc = torch.sigmoid(prv * inv_s), torch.sigmoid(nxt * inv_s)
    alpha = ((pc - nc + 1e-5) / (pc + 1e-5)).reshape(R, S).clip(0.0, 1.0) * pm.view(R, S)
    wts = alpha * torch.cumprod(torch.cat([torch.ones(R, 1), 1.0 - alpha + 1e-7], -1), -1)[:, :-1]
    wsum = wts.sum(-1, keepdim=True)
    color = (color_pts.view(R, S, 3) * wts[..., None]).sum(1) + background_rgb * (1.0 - wsum)
    depth = (mid * wts).sum(1, keepdim=True)
    cmask = ((nvalid.view(R, S) >= 2).float().sum(1, keepdim=True)) > 8
    return {"color": color, "depth": depth, "weights": wts, "gradients": grad.view(R, S, 3),
            "sdf": sdf.view(R, S), "inside": pm.view(R, S), "mid_z": mid, "color_mask": cmask,
            "weights_sum": wsum}


def render_rays(rays_o, rays_d, near, far, volume, occ, fmaps, imgs, w2cs, intrinsics, query_c2w,
                sd_sdf, sd_ren, variance, n_samples=64, n_importance=64, **kw):
    """SparseNeuSRenderer.render with perturb_overwrite=0 (reference sparse_neus_renderer.py:457-635)."""
    sample_dist = float((far - near) / n_samples)
    z = hierarchical_z(rays_o, rays_d, near, far, volume, occ, sd_sdf, n_samples, n_importance)
    out = render_fine(rays_o, rays_d, z, sample_dist, volume, occ, fmaps, imgs, w2cs, intrinsics,
                      query_c2w, sd_sdf, sd_ren, variance, **kw)
    out["z"] = z
    return out


# --------------------------------------------------------------------------------------
# B10  dense SDF grid + marching cubes (PyMCubes restatement, PARITY UNPINNED)
# --------------------------------------------------------------------------------------


def sdf_grid(volume, sd, resolution, chunk=65536):
    """u[x,y,z] = -sdf on linspace(-1,1,R)^3 (reference sparse_neus_renderer.py:882-905)."""
    lin = torch.linspace(-1.0, 1.0, resolution)
    pts = torch.stack(torch.meshgrid(lin, lin, lin, indexing="ij"), -1).reshape(-1, 3)
    out = torch.cat([sdf_query(p, volume, sd)[0] for p in pts.split(chunk)], 0)
    return (-out).reshape(resolution, resolution, resolution).numpy()


def marching_cubes(u, iso=0.0):
    """Classic marching cubes with one shared vertex per sign-changing lattice edge, linear
    interpolation in index units, float64 (PyMCubes semantics).

    Returns (vertices float64 [nv,3], triangles int64 [nt,3], case_index uint8 [X-1,Y-1,Z-1]).
    Vertex order: ascending (lattice point, axis).  A corner is inside when u > iso.
    """
    import sys
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "one-2-3-45_b200")
    if here not in sys.path:
        sys.path.insert(0, here)
    from o2345 import mc_tables as T
    _, tri_table, n_tri = T.tables()
    u = np.asarray(u, np.float64)
    X, Y, Z = u.shape
    ins = u > iso
    case = np.zeros((X - 1, Y - 1, Z - 1), np.uint8)
    for i, (dx, dy, dz) in enumerate(T.CORNERS):
        case |= (ins[dx:X - 1 + dx, dy:Y - 1 + dy, dz:Z - 1 + dz].astype(np.uint8) << i)
    vid = -np.ones((X, Y, Z, 3), np.int64)
    cross = np.zeros((X, Y, Z, 3), bool)
    cross[:-1, :, :, 0] = ins[:-1] != ins[1:]
    cross[:, :-1, :, 1] = ins[:, :-1] != ins[:, 1:]
    cross[:, :, :-1, 2] = ins[:, :, :-1] != ins[:, :, 1:]
    flat = np.nonzero(cross.reshape(-1))[0]
    vid.reshape(-1)[flat] = np.arange(flat.size)
    p = flat // 3
    ax = flat % 3
    px, py, pz = p // (Y * Z), (p // Z) % Y, p % Z
    qx, qy, qz = px + (ax == 0), py + (ax == 1), pz + (ax == 2)
    f0, f1 = u[px, py, pz], u[qx, qy, qz]
    t = (iso - f0) / (f1 - f0)
    verts = np.stack([px, py, pz], 1).astype(np.float64)
    verts[np.arange(flat.size), ax] += t
    tris = []
    cells = np.nonzero((case > 0) & (case < 255))
    cx, cy, cz = cells
    cc = case[cells]
    for k in range(5):
        sel = n_tri[cc] > k
        if not sel.any():
            break
        tri = []
        for j in range(3):
            e = tri_table[cc[sel], 3 * k + j].astype(np.int64)
            own = T.EDGE_OWNER[e]
            tri.append(vid[cx[sel] + own[:, 0], cy[sel] + own[:, 1], cz[sel] + own[:, 2], own[:, 3]])
        tris.append(np.stack(tri, 1))
    tris = np.concatenate(tris, 0) if tris else np.zeros((0, 3), np.int64)
    return verts, tris, case


def extract_geometry(volume, sd, resolution):
    """vertices in [-1,1]^3, triangles, u (reference sparse_neus_renderer.py:908-937)."""
    u = sdf_grid(volume, sd, resolution)
    v, t, _ = marching_cubes(u, 0.0)
    return v / (resolution - 1.0) * 2.0 - 1.0, t, u


def to_torch_state(sd_np):
    return {k: torch.from_numpy(np.asarray(v)).float() for k, v in sd_np.items()}
