"""Scene samples (row B0) and network assembly for the reconstruction path.

`load_sample` mirrors BlenderPerView.__getitem__ (reference data/One2345_eval_new_data.py:139-377)
for a folder written by run.py (pose.json, stage1_8/0.png, stage2_8/*.png); `synthetic_sample`
builds the same dict from seeded inputs.  `build_networks` mirrors Runner.__init__ (reference
exp_runner_generic_blender_val.py:93-151) for the lod-0 demo configuration.
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from . import synthetic as S
from .featurenet import FeatureNet
from .rendering_network import GeneralRenderingNetwork, SingleVarianceNetwork
from .sparse_sdf_network import SparseSdfNetwork
from .trainer_generic import GenericTrainer


from .checkpoints import Conf  # noqa: E402,F401


def build_networks(device, vol_dim=96, states=None, n_samples=64, n_importance=64, perturb=1.0, base_exp_dir=None,
                   variance_init=0.3, conf=None):
    """FeatureNet, SparseSdfNetwork, SingleVarianceNetwork, GeneralRenderingNetwork, GenericTrainer on `device`, assembled
    like Runner.__init__ (reference exp_runner_generic_blender_val.py:93-132).  With `conf` (a parsed
    confs/one2345_lod0_val_demo.conf) the constructor arguments come from it, exactly as the reference passes
    `**conf['model.sdf_network_lod0']` etc. -- including its 8-digit voxel_size 0.02105263; without it the same
    constants are built in (voxel_size = 2 / (D - 1) in full precision)."""
    fnet = FeatureNet()
    if conf is not None:
        if conf.get_int('model.num_lods') != 1:
            raise NotImplementedError("num_lods > 1 (the lod-1 refinement networks) is a 'next' row, SURVEY.md 8(f) item 3")
        sdf = SparseSdfNetwork(**conf['model.sdf_network_lod0'])
        var = SingleVarianceNetwork(**conf['model.variance_network'])
        rnet = GeneralRenderingNetwork(**conf['model.rendering_network'])
        tk = dict(conf['model.trainer'])
    else:
        sdf = SparseSdfNetwork(lod=0, ch_in=56, voxel_size=2.0 / (vol_dim - 1), vol_dims=[vol_dim] * 3, hidden_dim=128,
                               cost_type='variance_mean', d_pyramid_feature_compress=16, regnet_d_out=16,
                               num_sdf_layers=4, multires=6)
        var = SingleVarianceNetwork(variance_init)
        rnet = GeneralRenderingNetwork(in_geometry_feat_ch=16, in_rendering_feat_ch=56, anti_alias_pooling=True)
        tk = dict(n_samples_lod0=n_samples, n_importance_lod0=n_importance, n_samples_lod1=64, n_importance_lod1=64,
                  n_outside=0, perturb=perturb, alpha_type='div')
    if states is not None:
        load = lambda m, sd: m.load_state_dict({k: torch.as_tensor(np.asarray(v)) for k, v in sd.items()}, strict=False)
        for m, key in ((fnet, "pyramid_feature_network"), (sdf, "sdf_network_lod0"), (rnet, "rendering_network_lod0"),
                       (var, "variance_network_lod0")):
            res = load(m, states[key])
            assert not res.unexpected_keys, res.unexpected_keys
            assert all("num_batches_tracked" in k for k in res.missing_keys), res.missing_keys
    for m in (fnet, sdf, var, rnet):
        m.to(device)
        for p in m.parameters():
            p.requires_grad_(False)
    if conf is None:
        conf = Conf({"general": Conf({"base_exp_dir": base_exp_dir}), "model": Conf({"num_lods": 1})})
    trainer = GenericTrainer(None, fnet, None, sdf, None, var, None, rnet, None, tk["n_samples_lod0"], tk["n_importance_lod0"],
                             tk["n_samples_lod1"], tk["n_importance_lod1"], tk["n_outside"], tk["perturb"],
                             alpha_type=tk["alpha_type"], conf=conf, base_exp_dir=base_exp_dir)
    return trainer


def _sample_from(cams, imgs, device, H, W, pin=False):
    """Adds the batch dimension the reference's DataLoader adds and moves everything to `device`."""
    t = lambda x: x if torch.is_tensor(x) else torch.from_numpy(np.ascontiguousarray(x))     # device-resident views pass through
    rays_o, rays_v = S.query_rays(cams["query_intrinsic"], cams["query_c2w"], H, W)
    host = {
        "images": t(imgs[1:])[None], "query_image": t(imgs[0])[None], "w2cs": t(cams["w2cs"])[None],
        "c2ws": t(cams["c2ws"])[None], "intrinsics": t(cams["intrinsics"])[None],
        "affine_mats": t(cams["affine_mats"])[None], "query_c2w": t(cams["query_c2w"])[None],
        "query_w2c": t(cams["query_w2c"])[None], "query_near_far": t(cams["query_near_far"])[None],
        "scale_mat": t(cams["scale_mat"])[None], "trans_mat": t(cams["trans_mat"])[None],
        "partial_vol_origin": t(cams["partial_vol_origin"])[None], "img_wh": torch.tensor([[W, H]]),
    }
    rays = {"rays_o": t(rays_o)[None], "rays_v": t(rays_v)[None]}
    if pin:
        host = {k: (v if v.is_cuda else v.pin_memory()) for k, v in host.items()}
        rays = {k: v.pin_memory() for k, v in rays.items()}
    sample = {k: v.to(device, non_blocking=pin) for k, v in host.items()}
    sample["rays"] = {k: v.to(device, non_blocking=pin) for k, v in rays.items()}
    sample["batch_idx"], sample["meta"] = torch.tensor([0]), ["synthetic"]
    return sample, host, rays


def synthetic_sample(device, n_views=32, H=256, W=256, seed=1234, elev=60.0):
    """One scene with seeded images: 1 query view + n_views source views."""
    meta = S.pose_json(elev)
    k = np.array(meta["intrinsics"])
    k[:2] *= W / 256.0
    meta["intrinsics"] = k.tolist()
    cams = S.scene_cameras(meta, n_src=n_views, img_wh=(W, H))
    imgs = S.images(n_views + 1, H, W, seed=seed)
    return _sample_from(cams, imgs, device, H, W)[0]


def load_sample(folder, device):
    """Reads <folder>/pose.json, stage1_8/<first id>, stage2_8/<ids 8..39> like the reference dataset."""
    from PIL import Image
    meta = json.load(open(os.path.join(folder, "pose.json")))
    ids = list(meta["c2ws"].keys())

    def read(path):
        a = np.asarray(Image.open(path), np.float32) / 255.0
        a = a.transpose(2, 0, 1)
        if a.shape[0] == 4:
            a = a[:3] * a[-1:] + (1 - a[-1:])
        return a

    imgs = [read(os.path.join(folder, "stage1_8", ids[0]))]
    imgs += [read(os.path.join(folder, "stage2_8", ids[v])) for v in range(8, 40)]
    imgs = np.stack(imgs).astype(np.float32)
    H, W = imgs.shape[2:]
    cams = S.scene_cameras(meta, n_src=32, img_wh=(W, H))
    return _sample_from(cams, imgs, device, H, W)[0]


# --------------------------------------------------------------------------------------
# run.py end to end (reference run.py:79-119): image -> 8 + 32 generated views -> mesh
# --------------------------------------------------------------------------------------
def sample_from_views(stage1, stage2, pose, device, pin=False):
    """The batch dict of BlenderPerView built from in-memory views instead of PNG files."""
    ids = list(pose["c2ws"].keys())
    first = int(ids[0].split(".")[0])
    views = [stage1[first]] + [stage2[ids[v].split(".")[0]] for v in range(8, 40)]
    if torch.is_tensor(views[0]):      # uint8 [H, W, 3] on the device (generate_views(keep_on_device=True)): the same u8 / 255 there
        imgs = (torch.stack(views).to(torch.float32) / 255.0).permute(0, 3, 1, 2).contiguous()
    else:
        imgs = np.stack([(u8.astype(np.float32) / 255.0).transpose(2, 0, 1) for u8 in views]).astype(np.float32)
    H, W = imgs.shape[2:]
    cams = S.scene_cameras(pose, n_src=32, img_wh=(W, H))
    return _sample_from(cams, imgs, device, H, W, pin=pin)[0]


@torch.no_grad()
def image_to_mesh(zero123, trainer, input_u8, polar_angle=60, resolution=256, ddim_steps=75, stage2_steps=50, scale=3.0,
                  exp_dir=None, batched=True):
    """`python run.py --img_path X --half_precision` without SAM / elevation estimation: Zero123 stage 1 + stage 2
    (the reference's 10 DDIM sampler calls, run as two batched ones unless batched=False), camera set-up, cost volume, SDF grid, marching cubes, vertex colours.
    Returns dict(vertices, triangles, colors) as host numpy arrays (and writes mesh.ply when exp_dir is given)."""
    from .zero123 import generate_views
    dev = next(trainer.parameters()).device
    stage1, stage2, pose = generate_views(zero123, input_u8, polar_angle, ddim_steps, stage2_steps, scale, exp_dir, dev, batched=batched,
                                          keep_on_device=exp_dir is None)
    sample = sample_from_views(stage1, stage2, pose, dev)
    trainer.base_exp_dir = exp_dir
    return trainer(sample, mode="export_mesh", resolution=resolution)
