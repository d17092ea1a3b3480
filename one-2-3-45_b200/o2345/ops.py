"""Tensor-level wrappers over the C-ABI: torch owns device memory and streams, the library does the work.

Every function takes CUDA fp32 (or the stated integer) tensors, allocates the outputs and scratch the
C-ABI asks for, and enqueues on torch's current stream.  Nothing here computes on the host or with
PyTorch kernels (only allocation / zero-fill / pointer plumbing).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L
from . import mc_tables

_f32, _i32, _u8 = torch.float32, torch.int32, torch.uint8


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t, dtype=None):
    if t is None:
        return None
    if not t.is_cuda:
        raise L.O2345Error("expected a CUDA tensor (the o2345 kernels have no CPU path)")
    if t.device.index != torch.cuda.current_device():
        # the C-ABI launches on the current device's current stream: a tensor of another GPU would be read through a pointer that
        # is not valid there (wrap the call in `with torch.cuda.device(t.device)`, as one process per GPU does by construction)
        raise L.O2345Error(f"tensor on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}")
    if dtype is not None and t.dtype != dtype:
        raise L.O2345Error(f"expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise L.O2345Error("expected a contiguous tensor")
    return C.c_void_p(t.data_ptr())


def _f(t):
    return _p(t, _f32)


def cf32(t):
    """fp32 + contiguous view/copy of a CUDA tensor."""
    return t.detach().to(_f32).contiguous()


# ----------------------------------------------------------------------------- point sources
class PointSource:
    """Keeps the tensors alive that an o2345_points struct points to."""

    def __init__(self, struct, n, keep):
        self.struct, self.n, self._keep = struct, n, keep

    @staticmethod
    def explicit(pts):
        pts = cf32(pts).view(-1, 3)
        s = L.Points(mode=L.PTS_EXPLICIT, pts=pts.data_ptr())
        return PointSource(s, pts.shape[0], (pts,))

    @staticmethod
    def lattice(lin):
        lin = cf32(lin)
        R = lin.numel()
        return PointSource(L.Points(mode=L.PTS_LATTICE, lin=lin.data_ptr(), R=R), R ** 3, (lin,))

    @staticmethod
    def rays(rays_o, rays_d, z):
        rays_o, rays_d, z = cf32(rays_o), cf32(rays_d), cf32(z)
        R, S = z.shape
        s = L.Points(mode=L.PTS_RAYS, rays_o=rays_o.data_ptr(), rays_d=rays_d.data_ptr(), z=z.data_ptr(), S=S,
                     z_stride=S)
        return PointSource(s, R * S, (rays_o, rays_d, z))


# ----------------------------------------------------------------------------- SDF query (B8/B9)
def sdf_pack_weights(w0, b0, w1, b1, w2, b2):
    pack = torch.empty(L.SDF_PACK_FLOATS, dtype=_f32, device=w0.device)
    L.call("o2345_sdf_pack_weights", _f(cf32(w0)), _f(cf32(b0)), _f(cf32(w1)), _f(cf32(b1)), _f(cf32(w2)),
           _f(cf32(b2)), _f(pack), _stream())
    return pack


# SDF MLP kernel used when a call does not say otherwise: forward GEMMs on tensor cores with split-fp16 operands
# (fp32-grade values); L.SDF_FP32 selects the fp32 FMA kernel.
SDF_PRECISION = L.SDF_TC_SPLIT


def sdf_query(src: PointSource, vol_cl, pack, active=None, inactive_sdf=100.0, negate=False, want_feat=False,
              want_latent=False, want_grad=False, precision=None):
    """Returns dict(sdf [n,1], feat [n,127]?, latent [n,16]?, grad [n,3]?)."""
    precision = SDF_PRECISION if precision is None else precision
    n, dev = src.n, vol_cl.device
    D = vol_cl.shape[0]
    out = {"sdf": torch.empty(n, 1, dtype=_f32, device=dev)}
    if want_feat:
        out["feat"] = torch.empty(n, 127, dtype=_f32, device=dev)
    if want_latent:
        out["latent"] = torch.empty(n, 16, dtype=_f32, device=dev)
    if want_grad:
        out["grad"] = torch.empty(n, 3, dtype=_f32, device=dev)
    L.call("o2345_sdf_query", C.byref(src.struct), n, _f(vol_cl), D, _f(pack), _p(active, _u8),
           float(inactive_sdf), int(bool(negate)), int(precision), _f(out["sdf"]), _f(out.get("feat")), _f(out.get("latent")),
           _f(out.get("grad")), _stream())
    return out


# ----------------------------------------------------------------------------- cost volume (B3-B7)
def compact(flags):
    """flags uint8 [n] -> rows int32 [n] (first *count valid), index int32 [n], count int32 [1]."""
    n = flags.numel()
    dev = flags.device
    rows = torch.empty(n, dtype=_i32, device=dev)
    index = torch.empty(n, dtype=_i32, device=dev)
    count = torch.empty(1, dtype=_i32, device=dev)
    scratch = torch.empty(L.load().o2345_compact_scratch_ints(n), dtype=_i32, device=dev)
    L.call("o2345_compact", _p(flags, _u8), n, _p(rows, _i32), _p(index, _i32), _p(count, _i32), _p(scratch, _i32),
           _stream())
    return rows, index, count


def frustum_mask(proj, origin, voxel_size, D, sizeH, sizeW, min_views):
    V = proj.shape[0]
    bits = torch.empty(D ** 3, dtype=_i32, device=proj.device)
    keep = torch.empty(D ** 3, dtype=_u8, device=proj.device)
    L.call("o2345_frustum_mask", _f(proj), V, _f(origin), float(voxel_size), D, int(sizeH), int(sizeW),
           int(min_views), _p(bits, _i32), _p(keep, _u8), _stream())
    return bits, keep


def costvol_gather(feats_nhwc, proj, origin, voxel_size, D, sizeH, sizeW, rows, count, bits, max_rows):
    V, h, w, c = feats_nhwc.shape
    assert c == 16
    cost = torch.empty(max_rows, 32, dtype=_f32, device=proj.device)
    L.call("o2345_costvol_gather", _f(feats_nhwc), V, h, w, int(sizeH), int(sizeW), _f(proj), _f(origin),
           float(voxel_size), D, _p(rows, _i32), _p(count, _i32), max_rows, _p(bits, _i32), _f(cost), _stream())
    return cost


def dense_scatter(feat, rows, count, D, max_rows, want_cf=True):
    dev = feat.device
    vol_cl = torch.empty(D, D, D, 16, dtype=_f32, device=dev)
    vol_cf = torch.empty(1, 16, D, D, D, dtype=_f32, device=dev) if want_cf else None
    occ = torch.empty(1, 1, D, D, D, dtype=_f32, device=dev)
    L.call("o2345_dense_scatter", _f(feat), _p(rows, _i32), _p(count, _i32), max_rows, D, _f(vol_cl), _f(vol_cf),
           _f(occ), _stream())
    return vol_cl, vol_cf, occ


def occ_nearest(src: PointSource, occ):
    D = occ.shape[-1]
    out = torch.empty(src.n, dtype=_u8, device=occ.device)
    L.call("o2345_occ_nearest", C.byref(src.struct), src.n, _f(occ), D, _p(out, _u8), _stream())
    return out


# ----------------------------------------------------------------------------- sparse conv (B6)
class SparseLevel:
    """index lattice [E^3], row list, device-side count and a host-side upper bound of the rows."""

    def __init__(self, E, rows, index, count, max_rows):
        self.E, self.rows, self.index, self.count, self.max_rows = E, rows, index, count, max_rows


def sp_coarsen(level: SparseLevel) -> SparseLevel:
    Ec = level.E // 2 + 1
    dev = level.rows.device
    flags = torch.empty(Ec ** 3, dtype=_u8, device=dev)
    cmin = torch.empty(3, dtype=_i32, device=dev)
    L.call("o2345_sp_coarsen", _p(level.index, _i32), level.E, _p(level.rows, _i32), _p(level.count, _i32),
           level.max_rows, Ec, _p(flags, _u8), _p(cmin, _i32), _stream())
    rows, index, count = compact(flags)
    return SparseLevel(Ec, rows, index, count, Ec ** 3)


def sp_conv_bn_relu(x, lin: SparseLevel, lout: SparseLevel, mode, kernel, gamma, beta, skip=None, eps=1e-5):
    """One BasicSparse(De)ConvolutionBlock: conv -> BatchNorm(batch stats) -> ReLU (+ skip)."""
    cin, cout = kernel.shape[1], kernel.shape[2]
    dev = x.device
    raw = torch.empty(lout.max_rows, cout, dtype=_f32, device=dev)
    stats = torch.empty(2 * cout, dtype=torch.float64, device=dev)
    L.call("o2345_sp_conv", _f(x), _p(lin.index, _i32), lin.E, _p(lout.rows, _i32), _p(lout.count, _i32),
           lout.max_rows, lout.E, mode, _f(kernel), cin, cout, _f(raw), _p(stats, torch.float64), _stream())
    L.call("o2345_sp_bn_relu", _f(raw), _p(lout.count, _i32), lout.max_rows, cout, _p(stats, torch.float64),
           _f(gamma), _f(beta), float(eps), _f(skip), _f(raw), _stream())
    return raw


# ----------------------------------------------------------------------------- marching cubes (B10)
_MC_CACHE = {}


def _mc_tables(dev):
    key = str(dev)
    if key not in _MC_CACHE:
        _, tri, ntri = mc_tables.tables()
        _MC_CACHE[key] = (torch.from_numpy(tri.copy()).to(dev), torch.from_numpy(ntri.copy()).to(dev),
                          torch.from_numpy(mc_tables.EDGE_OWNER.astype(np.int8)).to(dev).contiguous())
    return _MC_CACHE[key]


def marching_cubes(u, iso=0.0):
    """u float32 [R,R,R] (device) -> verts float64 [nv,3] (index units), tris int32 [nt,3], cases uint8."""
    R = u.shape[0]
    dev = u.device
    u = cf32(u)
    tri, ntri, owner = _mc_tables(dev)
    cases = torch.empty((R - 1) ** 3, dtype=_u8, device=dev)
    cell_flags = torch.empty((R - 1) ** 3, dtype=_u8, device=dev)
    edge_flags = torch.empty(3 * R ** 3, dtype=_u8, device=dev)
    L.call("o2345_mc_classify", _f(u), R, float(iso), _p(cases, _u8), _p(cell_flags, _u8), _p(edge_flags, _u8),
           _stream())
    edges, vert_index, nv_d = compact(edge_flags)
    cells, _, nc_d = compact(cell_flags)
    nv, nc = int(nv_d.item()), int(nc_d.item())  # the mesh size has to reach the host anyway
    verts = torch.empty(nv, 3, dtype=torch.float64, device=dev)
    L.call("o2345_mc_vertices", _f(u), R, float(iso), _p(edges, _i32), _p(nv_d, _i32), nv,
           _p(verts, torch.float64), _stream())
    if nc == 0:
        return verts, torch.empty(0, 3, dtype=_i32, device=dev), cases.view(R - 1, R - 1, R - 1)
    offs = torch.empty(nc, dtype=_i32, device=dev)
    total = torch.empty(1, dtype=_i32, device=dev)
    scratch = torch.empty(L.load().o2345_scan_scratch_ints(nc), dtype=_i32, device=dev)
    L.call("o2345_mc_tri_offsets", _p(cases, _u8), _p(cells, _i32), _p(nc_d, _i32), nc, _p(ntri, _u8),
           _p(offs, _i32), _p(total, _i32), _p(scratch, _i32), _stream())
    nt = int(total.item())
    tris = torch.empty(nt, 3, dtype=_i32, device=dev)
    L.call("o2345_mc_triangles", _p(cases, _u8), R, _p(cells, _i32), _p(nc_d, _i32), nc, _p(offs, _i32),
           _p(tri, torch.int8), _p(ntri, _u8), _p(owner, torch.int8), _p(vert_index, _i32), _p(tris, _i32),
           _stream())
    return verts, tris, cases.view(R - 1, R - 1, R - 1)


# ----------------------------------------------------------------------------- FeatureNet (B1/B2)
def view_of(t, layout, c0=0):
    """View4 over tensor t laid out as 'nchw' or 'nhwc' (channel offset c0 for concatenation)."""
    if layout == "nchw":
        N, Cc, H, W = t.shape
        v = L.View4(ptr=t.data_ptr(), sn=Cc * H * W, sc=H * W, sh=W, sw=1, c0=c0)
    else:
        N, H, W, Cc = t.shape
        v = L.View4(ptr=t.data_ptr(), sn=H * W * Cc, sc=1, sh=W * Cc, sw=Cc, c0=c0)
    return v


def conv2d(x, weight, bias, stride, pad, want_stats):
    N, Cin, H, W = x.shape
    Cout, _, K, _ = weight.shape
    Ho, Wo = (H + 2 * pad - K) // stride + 1, (W + 2 * pad - K) // stride + 1
    out = torch.empty(N, Cout, Ho, Wo, dtype=_f32, device=x.device)
    stats = torch.empty(2 * Cout, dtype=torch.float64, device=x.device) if want_stats else None
    L.call("o2345_conv2d", _f(x), N, Cin, H, W, _f(weight), _f(bias), Cout, K, stride, pad, _f(out),
           _p(stats, torch.float64), _stream())
    return out, stats


def abn_apply(x, stats, gamma, beta, out_view, eps=1e-5, slope=0.01):
    N, Cc, H, W = x.shape
    L.call("o2345_abn_apply", _f(x), N, Cc, H, W, _p(stats, torch.float64), _f(gamma), _f(beta), float(eps),
           float(slope), C.byref(out_view), _stream())


def upsample_bilinear(x, factor, out_view, add=None):
    N, Cc, H, W = x.shape
    L.call("o2345_upsample_bilinear", _f(x), N, Cc, H, W, int(factor), _f(add), C.byref(out_view), _stream())


# ----------------------------------------------------------------------------- rendering (B11-B14)
def ray_upsample(rays_o, rays_d, z, sdf, inv_s, occ, u):
    R, S = z.shape
    n_new = u.numel()
    new_z = torch.empty(R, n_new, dtype=_f32, device=z.device)
    L.call("o2345_ray_upsample", _f(rays_o), _f(rays_d), R, _f(z), _f(sdf), S, float(inv_s), _f(occ),
           occ.shape[-1], _f(u), n_new, _f(new_z), _stream())
    return new_z


def ray_merge(z, sdf, new_z, new_sdf):
    R, S = z.shape
    n_new = new_z.shape[1]
    oz = torch.empty(R, S + n_new, dtype=_f32, device=z.device)
    osdf = torch.empty_like(oz)
    L.call("o2345_ray_merge", _f(z), _f(sdf), S, _f(new_z), _f(new_sdf), n_new, R, _f(oz), _f(osdf), _stream())
    return oz, osdf


def ray_midpoints(rays_o, rays_d, z, sample_dist, occ):
    R, S = z.shape
    mid = torch.empty_like(z)
    dists = torch.empty_like(z)
    active = torch.empty(R * S, dtype=_u8, device=z.device)
    L.call("o2345_ray_midpoints", _f(rays_o), _f(rays_d), R, _f(z), S, float(sample_dist), _f(occ), occ.shape[-1],
           _f(mid), _f(dists), _p(active, _u8), _stream())
    return mid, dists, active


class SourceViews:
    """Channel-last colour+feature maps and camera data of the source views (o2345_views)."""

    def __init__(self, maps_nhwc, proj34, centers, sizeW, sizeH):
        V, H, W, c = maps_nhwc.shape
        assert c == L.MAP_CH
        self.maps, self.proj, self.centers = maps_nhwc, cf32(proj34), cf32(centers)
        self.struct = L.Views(V=V, H=H, W=W, maps=self.maps.data_ptr(), proj=self.proj.data_ptr(),
                              centers=self.centers.data_ptr(), sizeW=float(sizeW), sizeH=float(sizeH))


def render_blend(src: PointSource, active, vol_cl, occ, views: SourceViews, rnet_pack, query_center=None, dirs=None,
                 precision=L.BLEND_TC_FP16):
    """precision: L.BLEND_TC_FP16 (tensor-core MLPs, fp16 operands / fp32 accumulate) or L.BLEND_FP32 (fp32 FMA)."""
    n, dev = src.n, vol_cl.device
    rgb = torch.empty(n, 3, dtype=_f32, device=dev)
    nvalid = torch.empty(n, dtype=_i32, device=dev)
    mode = 0 if dirs is None else 1
    L.call("o2345_render_blend", C.byref(src.struct), n, _p(active, _u8), _f(vol_cl), _f(occ), vol_cl.shape[0],
           C.byref(views.struct), mode, _f(query_center), _f(dirs), _f(rnet_pack), int(precision), _f(rgb), _p(nvalid, _i32),
           _stream())
    return rgb, nvalid


def ray_composite(rays_d, mid, dists, sdf, grad, color, active, nvalid, inv_s, ratio, background):
    R, S = mid.shape
    dev = mid.device
    out = {"color": torch.empty(R, 3, dtype=_f32, device=dev), "depth": torch.empty(R, 1, dtype=_f32, device=dev),
           "weights": torch.empty(R, S, dtype=_f32, device=dev), "cdf": torch.empty(R, S, dtype=_f32, device=dev),
           "alpha": torch.empty(R, S, dtype=_f32, device=dev),
           "weights_sum": torch.empty(R, 1, dtype=_f32, device=dev),
           "color_mask": torch.empty(R, 1, dtype=_u8, device=dev)}
    has_bg = background is not None
    L.call("o2345_ray_composite", _f(rays_d), R, S, _f(mid), _f(dists), _f(sdf), _f(grad), _f(color),
           _p(active, _u8), _p(nvalid, _i32), float(inv_s), float(ratio), int(has_bg),
           float(background if has_bg else 0.0), _f(out["color"]), _f(out["depth"]), _f(out["weights"]),
           _f(out["cdf"]), _f(out["alpha"]), _f(out["weights_sum"]), _p(out["color_mask"], _u8), _stream())
    return out
