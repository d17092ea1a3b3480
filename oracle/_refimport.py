"""Import the real One-2-3-45 reference on CPU (THIS CONTAINER ONLY).

TEST INFRASTRUCTURE.  `/root/reference` does not exist on the GPU box, so nothing that runs
there (tests -m gpu, smoke, bench) may import this module.  It is used by
`oracle/pin_against_reference.py` / `tests/golden/make_golden.py` to (a) check the
standalone restatement in `oracle/recon_oracle.py` against the reference's own Python and
(b) freeze golden vectors.

The reference needs third-party packages that are not installed:
torchsparse v1.4.0, inplace_abn, PyMCubes, icecream, trimesh, pyhocon, kornia.  They are
replaced by `sys.modules` stubs; the three that carry arithmetic (torchsparse, inplace_abn,
mcubes) are backed by the restatements in `recon_oracle.py` -- those stay "parity unpinned".
"""
from __future__ import annotations

import os
import sys
import types

REF = os.environ.get("O2345_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "reconstruction"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stubs():
    import torch
    import torch.nn as nn
    from . import recon_oracle as O

    class SparseTensor:
        """Minimal torchsparse.SparseTensor: feats [N,C], coords int32 [N,4] = (x,y,z,b)."""

        def __init__(self, feats, coords, stride=1):
            self.F, self.C, self.s = feats, coords, stride
            self.cmaps, self.kmaps = {}, {}

        def __add__(self, other):
            out = SparseTensor(self.F + other.F, self.C, self.s)
            out.cmaps, out.kmaps = self.cmaps, self.kmaps
            return out

    class Conv3d(nn.Module):
        def __init__(self, inc, outc, kernel_size=3, stride=1, dilation=1, bias=False, transposed=False):
            super().__init__()
            self.k, self.stride, self.transposed = kernel_size, stride, transposed
            self.kernel = nn.Parameter(torch.zeros(kernel_size ** 3, inc, outc))

        def forward(self, x):
            feats, coords, s, cm, km = O.torchsparse_conv3d(
                x.F, x.C, x.s, self.kernel, self.stride, self.transposed, x.cmaps, x.kmaps)
            out = SparseTensor(feats, coords, s)
            out.cmaps, out.kmaps = cm, km
            return out

    class BatchNorm(nn.BatchNorm1d):
        def forward(self, x):
            out = SparseTensor(super().forward(x.F), x.C, x.s)
            out.cmaps, out.kmaps = x.cmaps, x.kmaps
            return out

    class ReLU(nn.ReLU):
        def forward(self, x):
            out = SparseTensor(super().forward(x.F), x.C, x.s)
            out.cmaps, out.kmaps = x.cmaps, x.kmaps
            return out

    class InPlaceABN(nn.Module):
        def __init__(self, c, eps=1e-5, momentum=0.1, activation="leaky_relu", activation_param=0.01):
            super().__init__()
            self.weight = nn.Parameter(torch.ones(c))
            self.bias = nn.Parameter(torch.zeros(c))
            self.register_buffer("running_mean", torch.zeros(c))
            self.register_buffer("running_var", torch.ones(c))
            self.eps, self.slope = eps, activation_param

        def forward(self, x):
            return O.inplace_abn(x, self.weight, self.bias, self.eps, self.slope)

    ts = _mod("torchsparse", SparseTensor=SparseTensor, PointTensor=object)
    _mod("torchsparse.tensor", SparseTensor=SparseTensor, PointTensor=object)
    ts.nn = _mod("torchsparse.nn", Conv3d=Conv3d, BatchNorm=BatchNorm, ReLU=ReLU)
    ts.nn.functional = _mod("torchsparse.nn.functional")
    _mod("torchsparse.nn.utils", get_kernel_offsets=None)
    _mod("inplace_abn", InPlaceABN=InPlaceABN)
    _mod("mcubes", marching_cubes=lambda u, thr: O.marching_cubes(u, thr)[:2])
    _mod("icecream", ic=lambda *a, **k: None)
    _mod("trimesh")
    _mod("pyhocon", ConfigFactory=object, HOCONConverter=object)
    _mod("kornia", create_meshgrid=None)
    _mod("rembg")


def import_reference():
    """Returns a namespace of the reference modules used by the pin script."""
    if not available():
        raise RuntimeError("reference tree not present; this helper only works in the build container")
    install_stubs()
    for p in (os.path.join(REF, "reconstruction"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    ns = types.SimpleNamespace()
    import ops.back_project as back_project
    import ops.grid_sampler as grid_sampler
    import ops.generate_grids as generate_grids
    import models.sparse_sdf_network as sparse_sdf_network
    import models.sparse_neus_renderer as sparse_neus_renderer
    import models.rendering_network as rendering_network
    import models.projector as projector
    import models.render_utils as render_utils
    import models.featurenet as featurenet
    import models.fields as fields
    import models.rays as rays
    import tsparse.modules as tsparse_modules
    ns.__dict__.update(locals())
    return ns


class Conf(dict):
    """pyhocon-like shim for SparseNeuSRenderer(conf=...) (sparse_neus_renderer.py:46,62)."""

    def get_int(self, key, default=None):
        return int(self.get(key, default))
