"""SparseNeuSRenderer on the o2345 CUDA kernels.

Mirror of reference reconstruction/models/sparse_neus_renderer.py (constructor :31-42, render
:457-635, extract_fields / extract_geometry :882-937): same arguments, same result-dict keys.
What differs is where the work happens: depth samples, SDF values and colours stay on the device
from the first coarse sample to the composited pixel; marching cubes runs on the GPU.

Documented deviations (both have zero effect on colour / depth / weights):
  * `cat_z_vals` evaluates the SDF for new samples only `if torch.sum(pts_mask) > 1` over the
    whole ray CHUNK (reference :135); here a chunk with exactly one occupied new sample still gets
    its SDF evaluated.
  * `render_core` forces the first 100 samples valid when a chunk has no occupied sample (:222-223);
    those samples get alpha * mask = 0 anyway, so the branch is dropped.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from ._lib import inference_only
from . import _lib as L
from . import ops
from .featurenet import source_maps_channel_last
from .sparse_sdf_network import channel_last_volume


class SparseNeuSRenderer(nn.Module):
    def __init__(self, rendering_network_outside, sdf_network, variance_network, rendering_network, n_samples,
                 n_importance, n_outside, perturb, alpha_type='div', conf=None):
        super().__init__()
        if alpha_type != 'div' or n_outside != 0:
            raise NotImplementedError("only alpha_type='div' and n_outside=0 (the demo configuration) are accelerated")
        self.conf = conf
        self.base_exp_dir = conf['general.base_exp_dir'] if conf is not None else None
        self.rendering_network_outside = rendering_network_outside
        self.sdf_network, self.variance_network, self.rendering_network = sdf_network, variance_network, rendering_network
        self.n_samples, self.n_importance, self.n_outside = n_samples, n_importance, n_outside
        self.perturb, self.alpha_type = perturb, alpha_type
        self.if_fitted_rendering = False
        # view-blending MLPs: tensor-core kernel (fp16 operands, fp32 accumulate) by default; _lib.BLEND_FP32 selects the
        # fp32 FMA kernel whose operation order follows the reference (used by the tight parity tests)
        self.blend_precision = L.BLEND_TC_FP16
        self._views_key, self._views = None, None
        self._u = {}

    # ------------------------------------------------------------------ helpers
    def _source_views(self, feature_maps, color_maps, w2cs, intrinsics, img_wh):
        """Channel-last source maps + projection matrices, rebuilt whenever any input is a different tensor OBJECT or has
        been written to.  The cache holds strong references to the inputs it was built from: an address can only be
        re-used by the caching allocator after the tensor is gone, so identity + version cannot alias a new scene
        (keying on data_ptr alone could: a second image in the same process may land on the first one's address)."""
        ins = (feature_maps, color_maps, w2cs, intrinsics)
        key = self._views_key
        same = key is not None and all(a is b for a, b in zip(key[0], ins)) and key[1] == tuple(t._version for t in ins) \
            and key[2] == (float(img_wh[0]), float(img_wh[1]))
        if not same:
            maps = source_maps_channel_last(feature_maps, color_maps)
            w2cs_f, intr = w2cs.float(), intrinsics.float()
            proj = torch.matmul(intr, w2cs_f[:, :3, :])
            centers = torch.inverse(w2cs_f)[:, :3, 3]
            self._views = ops.SourceViews(maps, proj, centers, float(img_wh[0]), float(img_wh[1]))
            self._views_key = (ins, tuple(t._version for t in ins), (float(img_wh[0]), float(img_wh[1])))
        return self._views

    def _u_table(self, n, dev):
        k = (n, str(dev))
        if k not in self._u:
            self._u[k] = torch.linspace(0. + 0.5 / n, 1. - 0.5 / n, steps=n).to(dev)
        return self._u[k]

    # ------------------------------------------------------------------ B13 + B14
    @inference_only
    def render(self, rays_o, rays_d, near, far, sdf_network, rendering_network, perturb_overwrite=-1,
               background_rgb=None, alpha_inter_ratio=0.0, lod=None, conditional_volume=None,
               conditional_valid_mask_volume=None, feature_maps=None, color_maps=None, w2cs=None, intrinsics=None,
               img_wh=None, query_c2w=None, if_general_rendering=True, if_render_with_grad=True, img_index=None,
               rays_uv=None, pre_sample=False, bg_ratio=0.0):
        if bg_ratio != 0.0 or pre_sample or not if_general_rendering:
            raise NotImplementedError("bg_ratio / pre_sample / fitted rendering are training-time options")
        dev = rays_o.device
        rays_o, rays_d = ops.cf32(rays_o), ops.cf32(rays_d)
        R = rays_o.shape[0]
        n_s, n_i = self.n_samples, self.n_importance
        near_t = near if torch.is_tensor(near) else torch.tensor([near], device=dev)
        far_t = far if torch.is_tensor(far) else torch.tensor([far], device=dev)
        near_t, far_t = near_t.to(dev).float(), far_t.to(dev).float()
        sample_dist = ((far_t - near_t) / n_s).mean().item()
        z = near_t + (far_t - near_t) * torch.linspace(0.0, 1.0, n_s).to(dev)[None, :]
        if z.shape[0] == 1:
            z = z.repeat(R, 1)
        perturb = self.perturb if perturb_overwrite < 0 else perturb_overwrite
        if perturb > 0:  # stratified jitter, same draws as the reference (:508-515)
            mids = .5 * (z[..., 1:] + z[..., :-1])
            upper = torch.cat([mids, z[..., -1:]], -1)
            lower = torch.cat([z[..., :1], mids], -1)
            z = lower + (upper - lower) * torch.rand(z.shape).to(dev)
        z = z.contiguous()
        vol_cl = channel_last_volume(conditional_volume)
        occ = ops.cf32(conditional_valid_mask_volume)
        pack = sdf_network.sdf_layer.packed()

        if n_i > 0:
            sdf = ops.sdf_query(ops.PointSource.rays(rays_o, rays_d, z), vol_cl, pack)["sdf"].view(R, n_s)
            n_steps = 4
            u = self._u_table(n_i // n_steps, dev)
            for i in range(n_steps):
                new_z = ops.ray_upsample(rays_o, rays_d, z, sdf, 64 * 2 ** i, occ, u)
                src = ops.PointSource.rays(rays_o, rays_d, new_z)
                act = ops.occ_nearest(src, occ)
                new_sdf = ops.sdf_query(src, vol_cl, pack, active=act, inactive_sdf=100.0)["sdf"].view(R, -1)
                z, sdf = ops.ray_merge(z, sdf, new_z, new_sdf)
        S = z.shape[1]

        mid, dists, active = ops.ray_midpoints(rays_o, rays_d, z, sample_dist, occ)
        src = ops.PointSource.rays(rays_o, rays_d, mid)
        q = ops.sdf_query(src, vol_cl, pack, active=active, inactive_sdf=100.0, want_grad=True)
        views = self._source_views(feature_maps, color_maps, w2cs, intrinsics, img_wh)
        qc = ops.cf32(query_c2w.reshape(-1, 4, 4)[0, :3, 3])
        color_pts, nvalid = ops.render_blend(src, active, vol_cl, occ, views, rendering_network.packed(), query_center=qc,
                                             precision=self.blend_precision)
        inv_s = self.variance_network.inv_s()
        bg = None if background_rgb is None else float(background_rgb)
        comp = ops.ray_composite(rays_d, mid, dists, q["sdf"], q["grad"], color_pts, active, nvalid, inv_s,
                                 float(alpha_inter_ratio), bg)

        weights, depth = comp["weights"], comp["depth"]
        pts_mask = active.view(R, S).float()
        gradients = q["grad"].view(R, S, 3)
        gerr = (torch.linalg.norm(gradients, ord=2, dim=-1) - 1.0) ** 2
        gradient_error = (pts_mask * gerr).sum() / (pts_mask.sum() + 1e-5)
        pts_random = torch.rand([1024, 3]).float().to(dev) * 2 - 1
        sdf_random = ops.sdf_query(ops.PointSource.explicit(pts_random), vol_cl, pack)["sdf"]
        color_mask = comp["color_mask"].bool()
        return {
            'depth': depth,
            'color_fine': comp["color"],
            'color_fine_mask': color_mask,
            'color_outside': None,
            'color_outside_mask': None,
            'color_mlp': None,
            'color_mlp_mask': None,
            'variance': torch.tensor(1.0 / inv_s, device=dev),
            'cdf_fine': comp["cdf"],
            'depth_variance': ((mid - depth) ** 2 * weights).sum(dim=-1, keepdim=True),
            'weights_sum': comp["weights_sum"],
            'weights_max': torch.max(weights, dim=-1, keepdim=True)[0],
            'alpha_sum': comp["alpha"].sum(dim=-1, keepdim=True).mean(),
            'alpha_mean': comp["alpha"].mean(),
            'gradients': gradients,
            'weights': weights,
            'gradient_error_fine': gradient_error,
            'inside_sphere': pts_mask,
            'sdf': q["sdf"],
            'sdf_random': sdf_random,
            'blended_color_patch': None,
            'blended_color_patch_mask': None,
            'weights_sum_fg': comp["weights_sum"],
            'z_vals': z,
            'mid_z_vals': mid,
        }

    # ------------------------------------------------------------------ B10
    @torch.no_grad()
    def extract_fields(self, bound_min, bound_max, resolution, query_func, device, **kwargs):
        """u[x,y,z] = -sdf on the lattice, returned as a DEVICE tensor [R,R,R] (the reference copies 64^3
        chunks to host numpy, :901-904).  `query_func` is ignored: the lattice mode of the SDF kernel is used."""
        bmin = [float(v) for v in bound_min]
        bmax = [float(v) for v in bound_max]
        if len(set(bmin)) != 1 or len(set(bmax)) != 1:
            raise NotImplementedError("cubic bounds only")
        lin = torch.linspace(bmin[0], bmax[0], resolution).to(device)
        net = self.sdf_network
        out = ops.sdf_query(ops.PointSource.lattice(lin), channel_last_volume(kwargs["conditional_volume"]),
                            net.sdf_layer.packed(), negate=True)
        return out["sdf"].view(resolution, resolution, resolution)

    @inference_only
    def extract_geometry(self, sdf_network, bound_min, bound_max, resolution, threshold, device, occupancy_mask=None,
                         **kwargs):
        """-> (vertices float64 numpy [nv,3] in world units, triangles int numpy [nt,3], u device tensor)."""
        if occupancy_mask is not None:
            raise NotImplementedError("occupancy_mask is only used by the lod-1 path")
        prev, self.sdf_network = self.sdf_network, sdf_network
        try:
            u = self.extract_fields(bound_min, bound_max, resolution, None, device, **kwargs)
        finally:
            self.sdf_network = prev
        verts, tris, cases = ops.marching_cubes(u, float(threshold))
        self._last_cases = cases
        # marching cubes emits one vertex per sign-changing lattice edge: two vertices can only coincide at a lattice point.
        # Those within 1e-4 voxels of one are the only candidates of trimesh's vertex merge in the mesh tail.
        self.mc_lattice_candidates = (torch.nonzero((verts - verts.round()).abs().amax(dim=1) < 1e-4)[:, 0].cpu().numpy()
                                      if verts.numel() else np.zeros(0, np.int64))
        b_max = np.asarray([float(v) for v in bound_max])
        b_min = np.asarray([float(v) for v in bound_min])
        vertices = verts.cpu().numpy() / (resolution - 1.0) * (b_max - b_min)[None, :] + b_min[None, :]
        return vertices, tris.cpu().numpy(), u

    # ------------------------------------------------------------------ B11 + B12 for mesh vertices
    @inference_only
    def blend_points(self, pts, sdf_network, rendering_network, conditional_volume, conditional_valid_mask_volume,
                     feature_maps, color_maps, w2cs, intrinsics, img_wh):
        """Vertex colours: Projector.compute_view_independent + rendering network (reference
        projector.py:231-425, trainer_generic.py:1338-1361).  -> rgb [n,3], normals [n,3]."""
        vol_cl = channel_last_volume(conditional_volume)
        occ = ops.cf32(conditional_valid_mask_volume)
        src = ops.PointSource.explicit(pts)
        g = ops.sdf_query(src, vol_cl, sdf_network.sdf_layer.packed(), want_grad=True)["grad"]
        normals = torch.nn.functional.normalize(g, p=2, dim=-1, eps=1e-6)
        views = self._source_views(feature_maps, color_maps, w2cs, intrinsics, img_wh)
        rgb, _ = ops.render_blend(src, None, vol_cl, occ, views, rendering_network.packed(), dirs=normals.contiguous(),
                                  precision=self.blend_precision)
        return rgb, normals
