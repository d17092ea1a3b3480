"""Shared mini configuration of the parity tests (same inputs as oracle/pin_against_reference.py)."""
import numpy as np
import torch

from o2345 import synthetic as S
from oracle import recon_oracle as O
from oracle.pin_against_reference import MINI, mini_points, mini_scene  # noqa: F401


def t(x):
    return torch.from_numpy(np.asarray(x)).float()


def states_torch(seed=0):
    return {k: O.to_torch_state(v) for k, v in S.all_states(seed).items()}


class OracleMini:
    """Everything the oracle produces at the mini configuration, computed once per session."""

    def __init__(self):
        self.st = states_torch(0)
        self.cams, imgs = mini_scene()
        self.imgs = t(imgs)
        self.proj, self.origin = t(self.cams["affine_mats"]), t(self.cams["partial_vol_origin"])
        D, H, W = MINI["D"], MINI["H"], MINI["W"]
        self.voxel = 2.0 / (D - 1)
        self.fmaps = O.pyramid_feature_maps(self.imgs, self.st["pyramid_feature_network"])
        self.cv = O.conditional_volume(self.fmaps, self.origin, self.proj, self.st["sdf_network_lod0"], D, self.voxel, H, W)
        self.volume, self.occ = self.cv["dense"], self.cv["occ"]
        self.pts = t(mini_points(MINI["n_pts"]))
        ro, rv = S.query_rays(self.cams["query_intrinsic"], self.cams["query_c2w"], H, W)
        sel = np.linspace(0, H * W - 1, MINI["n_rays"]).astype(np.int64)
        self.rays_o, self.rays_d = t(ro[sel]), t(rv[sel])
        self.near, self.far = t(self.cams["query_near_far"][:1]), t(self.cams["query_near_far"][1:])
        self.w2cs, self.intr = t(self.cams["w2cs"]), t(self.cams["intrinsics"])
        self.qc2w = t(self.cams["query_c2w"])[None]
        self.verts = t(np.random.default_rng(9).uniform(-0.7, 0.7, size=(MINI["n_verts"], 3)).astype(np.float32))
