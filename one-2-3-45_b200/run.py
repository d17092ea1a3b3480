"""`python run.py --img_path P [--gpu_idx N] [--half_precision] [--mesh_resolution R] [--output_format .ply]`

Command-line mirror of the reference's run.py:99-119 for the two accelerated paths: Zero123 stage 1 + stage 2
(8 + 32 views, DDIM 75 / 50 steps, CFG 3) and the cost-volume reconstruction, writing the same artefacts under
./exp/<shape>/ (stage1_8/*.png, stage2_8/*.png, pose.json, mesh.ply).

Not built: SAM / rembg foreground extraction -- the input must already be a segmented object on a
plain (or transparent) background -- and the LoFTR elevation search (`--polar_angle`, default 60).  Checkpoints: without
`--zero123_ckpt` / `--recon_ckpt` the seeded synthetic weights of o2345.synthetic are used (there is no network access to
fetch the released ones); with them, the reference's own files load through load_state_dict.
"""
import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def load_input(path):
    """256 x 256 RGB uint8 on white, like the output of the reference's preprocess() (utils/zero123_utils.py:180-202)."""
    from PIL import Image
    im = Image.open(path)
    if im.mode == "RGBA":
        bg = Image.new("RGBA", im.size, (255, 255, 255, 255))
        im = Image.alpha_composite(bg, im)
    return np.asarray(im.convert("RGB").resize((256, 256), Image.LANCZOS), np.uint8)


def main(argv=None):
    ap = argparse.ArgumentParser(description="single image -> textured mesh on the o2345 (sm_100a) kernels")
    ap.add_argument('--img_path', type=str, default="./demo/demo_examples/01_wild_hydrant.png", help='Path to the input image')
    ap.add_argument('--gpu_idx', type=int, default=0, help='GPU index')
    ap.add_argument('--half_precision', action='store_true', help='accepted for compatibility: the UNet / VAE kernels are fp16')
    ap.add_argument('--mesh_resolution', type=int, default=256, help='Mesh resolution')
    ap.add_argument('--output_format', type=str, default=".ply", help='Output format: .ply (.obj / .glb need trimesh: not built)')
    ap.add_argument('--polar_angle', type=float, default=60.0, help='elevation of the input view in degrees (not estimated)')
    ap.add_argument('--zero123_ckpt', type=str, default=None, help='zero123-xl.ckpt (state_dict); default: seeded synthetic weights')
    ap.add_argument('--recon_ckpt', type=str, default=None, help='reconstruction checkpoint (ckpt_*.pth); default: seeded synthetic weights')
    args = ap.parse_args(argv)
    if not torch.cuda.is_available():
        raise SystemExit("run.py needs a CUDA device: the o2345 path has no CPU fallback")
    from o2345 import synthetic as S
    from o2345.pipeline import build_networks, image_to_mesh
    from o2345.zero123 import LatentDiffusion, build_zero123
    dev = torch.device("cuda", args.gpu_idx)
    torch.cuda.set_device(dev)

    if args.zero123_ckpt:
        sd = torch.load(args.zero123_ckpt, map_location="cpu")
        sd = sd.get("state_dict", sd)
        if any(k.startswith("cond_stage_model.") for k in sd):
            print("note: the CLIP image tower (cond_stage_model.*) is not built; a fixed embedding stands in for it", file=sys.stderr)
        model = LatentDiffusion()
        res = model.load_state_dict({k: v for k, v in sd.items() if not k.startswith(("cond_stage_model.", "model_ema."))},
                                    strict=False)
        print(f"zero123 checkpoint: {len(res.missing_keys)} missing / {len(res.unexpected_keys)} unexpected keys", file=sys.stderr)
        model = model.requires_grad_(False).to(dev)
    else:
        print("no --zero123_ckpt: seeded synthetic Zero123 weights (the generated views are noise-like)", file=sys.stderr)
        model = build_zero123(dev, seed=0)
    model = model.half()

    states = S.all_states(0)
    if args.recon_ckpt:
        ck = torch.load(args.recon_ckpt, map_location="cpu")
        states = {"pyramid_feature_network": ck.get("pyramid_feature_network", ck.get("pyramid_feature_network_lod0")),
                  "sdf_network_lod0": ck["sdf_network_lod0"], "rendering_network_lod0": ck["rendering_network_lod0"],
                  "variance_network_lod0": ck["variance_network_lod0"]}
    shape_id = os.path.basename(args.img_path).split('.')[0]
    shape_dir = os.path.join("exp", shape_id)
    os.makedirs(shape_dir, exist_ok=True)
    trainer = build_networks(dev, vol_dim=96, states=states, perturb=0.0, base_exp_dir=shape_dir)

    mesh = image_to_mesh(model, trainer, load_input(args.img_path), polar_angle=args.polar_angle,
                         resolution=args.mesh_resolution, exp_dir=shape_dir)
    ply = os.path.join(shape_dir, "mesh.ply")
    if args.output_format != ".ply":
        print("Invalid output format for this build (only .ply is written)", file=sys.stderr)
    print(f"{len(mesh['vertices'])} vertices, {len(mesh['triangles'])} triangles")
    print("Mesh saved to:", ply)
    return ply


if __name__ == "__main__":
    main()
