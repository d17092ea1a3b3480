"""Pin oracle/vae_oracle.py against the reference's own Decoder / Encoder (build container only) and freeze
tests/golden/vae_mini.npz.   python -m oracle.pin_vae_against_reference"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "one-2-3-45_b200"))
sys.path.insert(0, ROOT)
from o2345 import synthetic as S  # noqa: E402
from oracle import vae_oracle as VO  # noqa: E402


def vae_inputs():
    g = np.random.default_rng(4)
    z = g.standard_normal((1, 4, 32, 32), dtype=np.float32)
    x = g.uniform(-1, 1, (1, 3, 256, 256)).astype(np.float32)
    return z, x


def main():
    for name in ("matplotlib", "matplotlib.pyplot", "omegaconf", "omegaconf.listconfig", "taming", "kornia", "clip", "pytorch_lightning"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, "/root/reference")
    from ldm.modules.diffusionmodules.model import Decoder, Encoder
    sd = {k: torch.from_numpy(v) for k, v in S.vae_state().items()}
    dd = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4],
              num_res_blocks=2, attn_resolutions=[], dropout=0.0)
    dec, enc = Decoder(**dd), Encoder(**dd)
    dec.load_state_dict({k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")})
    enc.load_state_dict({k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")})
    z, x = map(torch.from_numpy, vae_inputs())
    with torch.no_grad():
        d_ref = dec(torch.nn.functional.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"]))
        m_ref = torch.nn.functional.conv2d(enc(x), sd["quant_conv.weight"], sd["quant_conv.bias"])
        e1 = float((d_ref - VO.decode(sd, z)).abs().max())
        e2 = float((m_ref - VO.encode_moments(sd, x)).abs().max())
    print(f"A6 decode max|ref-oracle| = {e1:.2e}   A7 encode moments = {e2:.2e}")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "vae_mini.npz"), dec=d_ref.numpy()[:, :, ::4, ::4], moments=m_ref.numpy())
    return 0 if max(e1, e2) < 1e-4 else 1


if __name__ == "__main__":
    sys.exit(main())
