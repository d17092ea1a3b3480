"""GenericTrainer (inference subset: mode='val' and mode='export_mesh') on the o2345 kernels.

Mirror of reference reconstruction/models/trainer_generic.py: constructor :18-125, forward dispatch
:1052-1103, val_step :359-545, export_mesh_step :827-979, obtain_pyramid_feature_maps :1104-1125,
validate_colored_mesh :1309-1380.  Training (`train_step`, losses) stays with the reference.
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch
import torch.nn as nn

from .featurenet import obtain_pyramid_feature_maps
from .sparse_neus_renderer import SparseNeuSRenderer


from .mesh_io import merge_vertices, write_ply  # noqa: E402,F401  (write_ply re-exported: tests and tools import it from here)


class GenericTrainer(nn.Module):
    def __init__(self, rendering_network_outside, pyramid_feature_network_lod0, pyramid_feature_network_lod1,
                 sdf_network_lod0, sdf_network_lod1, variance_network_lod0, variance_network_lod1,
                 rendering_network_lod0, rendering_network_lod1, n_samples_lod0, n_importance_lod0, n_samples_lod1,
                 n_importance_lod1, n_outside, perturb, alpha_type='div', conf=None, timestamp="", mode='train',
                 base_exp_dir=None):
        super().__init__()
        self.conf, self.timestamp, self.base_exp_dir = conf, timestamp, base_exp_dir
        self.rendering_network_outside = rendering_network_outside
        self.pyramid_feature_network_geometry_lod0 = pyramid_feature_network_lod0
        self.sdf_network_lod0 = sdf_network_lod0
        self.variance_network_lod0 = variance_network_lod0
        self.rendering_network_lod0 = rendering_network_lod0
        self.n_samples_lod0, self.n_importance_lod0 = n_samples_lod0, n_importance_lod0
        self.n_outside, self.perturb, self.alpha_type = n_outside, perturb, alpha_type
        self.num_lods = 1
        if sdf_network_lod1 is not None:
            raise NotImplementedError("num_lods > 1 is a 'next' row (SURVEY.md 8(f) item 3)")
        self.sdf_renderer_lod0 = SparseNeuSRenderer(rendering_network_outside, sdf_network_lod0, variance_network_lod0,
                                                    rendering_network_lod0, n_samples_lod0, n_importance_lod0,
                                                    n_outside, perturb, alpha_type='div', conf=conf)
        self.val_mesh_freq = 1

    def obtain_pyramid_feature_maps(self, imgs, lod=0):
        return obtain_pyramid_feature_maps(self.pyramid_feature_network_geometry_lod0, imgs)

    def forward(self, sample, perturb_overwrite=-1, background_rgb=None, alpha_inter_ratio_lod0=0.0,
                alpha_inter_ratio_lod1=0.0, iter_step=0, mode='train', save_vis=False, resolution=360):
        if mode == 'val':
            return self.val_step(sample, perturb_overwrite=perturb_overwrite, background_rgb=background_rgb,
                                 alpha_inter_ratio_lod0=alpha_inter_ratio_lod0, iter_step=iter_step, save_vis=save_vis)
        if mode == 'export_mesh':
            return self.export_mesh_step(sample, iter_step=iter_step, save_vis=save_vis, resolution=resolution)
        raise NotImplementedError(f"mode={mode!r}: only 'val' and 'export_mesh' run on the o2345 path")

    # ------------------------------------------------------------------ shared front end
    @torch.no_grad()
    def _conditional_features(self, sample):
        sizeW, sizeH = int(sample['img_wh'][0][0]), int(sample['img_wh'][0][1])
        imgs = sample['images'][0]
        fmaps = self.obtain_pyramid_feature_maps(imgs, lod=0)
        cond = self.sdf_network_lod0.get_conditional_volume(
            feature_maps=fmaps[None], partial_vol_origin=sample['partial_vol_origin'],
            proj_mats=sample['affine_mats'], sizeH=sizeH, sizeW=sizeW, lod=0)
        return imgs, fmaps, cond, sizeW, sizeH

    # ------------------------------------------------------------------ mode='val'
    @torch.no_grad()
    def val_step(self, sample, perturb_overwrite=-1, background_rgb=None, alpha_inter_ratio_lod0=0.0,
                 alpha_inter_ratio_lod1=0.0, iter_step=0, chunk_size=512, save_vis=False):
        """Renders the query view.  Returns dict(color [H*W,3], depth [H*W,1], normal [H*W,3]) as numpy,
        like the per-chunk host copies of the reference (:526-543) but with one copy at the end.
        `chunk_size` rays are marched per launch group (the reference uses 512; larger is faster)."""
        imgs, fmaps, cond, sizeW, sizeH = self._conditional_features(sample)
        vol, occ = cond['dense_volume_scale0'], cond['valid_mask_volume_scale0']
        near, far = sample['query_near_far'][0, :1], sample['query_near_far'][0, 1:]
        rays_o = sample['rays']['rays_o'][0].reshape(-1, 3)
        rays_d = sample['rays']['rays_v'][0].reshape(-1, 3)
        colors, depths, normals = [], [], []
        for ro, rd in zip(rays_o.split(chunk_size), rays_d.split(chunk_size)):
            out = self.sdf_renderer_lod0.render(
                ro, rd, near, far, self.sdf_network_lod0, self.rendering_network_lod0,
                perturb_overwrite=perturb_overwrite, background_rgb=background_rgb,
                alpha_inter_ratio=alpha_inter_ratio_lod0, lod=0, conditional_volume=vol,
                conditional_valid_mask_volume=occ, feature_maps=fmaps, color_maps=imgs, w2cs=sample['w2cs'][0],
                intrinsics=sample['intrinsics'][0], img_wh=[sizeW, sizeH], query_c2w=sample['query_c2w'],
                if_render_with_grad=False)
            colors.append(out['color_fine'])
            depths.append(out['depth'])
            normals.append((out['gradients'] * out['weights'][:, :, None] * out['inside_sphere'][..., None]).sum(dim=1))
        return {"color": torch.cat(colors).cpu().numpy(), "depth": torch.cat(depths).cpu().numpy(),
                "normal": torch.cat(normals).cpu().numpy()}

    # ------------------------------------------------------------------ mode='export_mesh'
    @torch.no_grad()
    def export_mesh_step(self, sample, iter_step=0, chunk_size=512, resolution=360, save_vis=False):
        imgs, fmaps, cond, sizeW, sizeH = self._conditional_features(sample)
        return self.validate_colored_mesh(
            density_or_sdf_network=self.sdf_network_lod0,
            func_extract_geometry=self.sdf_renderer_lod0.extract_geometry, resolution=resolution,
            conditional_volume=cond['dense_volume_scale0'],
            conditional_valid_mask_volume=cond['valid_mask_volume_scale0'], feature_maps=fmaps, color_maps=imgs,
            w2cs=sample['w2cs'][0], intrinsics=sample['intrinsics'][0],
            rendering_network=self.rendering_network_lod0, lod=0, threshold=0, query_c2w=sample['query_c2w'],
            scale_mat=sample['scale_mat'], trans_mat=sample['trans_mat'], img_wh=[sizeW, sizeH])

    @torch.no_grad()
    def validate_colored_mesh(self, density_or_sdf_network, func_extract_geometry, world_space=True, resolution=360,
                              threshold=0.0, mode='val', conditional_volume=None, conditional_valid_mask_volume=None,
                              feature_maps=None, color_maps=None, w2cs=None, target_candidate_w2cs=None,
                              intrinsics=None, rendering_network=None, rendering_projector=None, query_c2w=None,
                              lod=None, occupancy_mask=None, bound_min=[-1, -1, -1], bound_max=[1, 1, 1], meta='',
                              iter_step=0, scale_mat=None, trans_mat=None, img_wh=(256, 256)):
        bmin = torch.tensor(bound_min, dtype=torch.float32)
        bmax = torch.tensor(bound_max, dtype=torch.float32)
        vertices, triangles, fields = func_extract_geometry(
            density_or_sdf_network, bmin, bmax, resolution=resolution, threshold=threshold,
            device=conditional_volume.device, conditional_volume=conditional_volume, lod=lod,
            occupancy_mask=occupancy_mask)
        vt = torch.tensor(vertices).to(conditional_volume)
        rgb, _ = self.sdf_renderer_lod0.blend_points(vt, density_or_sdf_network, rendering_network, conditional_volume,
                                                     conditional_valid_mask_volume, feature_maps, color_maps, w2cs,
                                                     intrinsics, img_wh)
        if scale_mat is not None:
            sm = scale_mat.cpu().numpy()
            vertices = vertices * sm[0][0, 0] + sm[0][:3, 3][None]
        if trans_mat is not None:
            tm = trans_mat.cpu().numpy().reshape(-1, 4, 4)[0]
            vh = np.concatenate([vertices, np.ones_like(vertices[:, :1])], axis=1)
            vertices = (vh @ tm.T)[:, :3]
        colors = (rgb.cpu() * 255).numpy().astype(np.uint8)
        # trimesh.Trimesh(vertices, triangles, vertex_colors=...) with its default process=True merges coincident vertices
        # before the export (reference :1374-1380).  Marching-cubes vertices can only coincide on lattice points: the renderer
        # listed the vertices that sit on one (on the device), and only those are compared.
        vertices, triangles, colors = merge_vertices(vertices, triangles, colors,
                                                     candidates=getattr(self.sdf_renderer_lod0, "mc_lattice_candidates", None))
        if self.base_exp_dir is not None:
            os.makedirs(self.base_exp_dir, exist_ok=True)
            write_ply(os.path.join(self.base_exp_dir, 'mesh.ply'), vertices, triangles, colors)
        return {"vertices": vertices, "triangles": triangles, "colors": colors, "fields": fields}
