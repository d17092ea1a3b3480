"""`python exp_runner_generic_blender_val.py --specific_dataset_name <exp_dir> --mode export_mesh --conf C --resolution R`

Command-line mirror of the reference's reconstruction/exp_runner_generic_blender_val.py:596-640 for the two inference
modes of the lod-0 demo configuration:
  export_mesh   <exp_dir>/{pose.json, stage1_8, stage2_8} -> <exp_dir>/mesh.ply   (what run.py's reconstruct() shells out to)
  val           volume-renders the query view -> <exp_dir>/val_color.png, val_depth.npy, val_normal.npy
`--conf` is parsed (o2345/checkpoints.py: the HOCON subset the reference's confs use; pyhocon is not needed) and supplies
`model.sdf_network_lod0` (voxel_size, vol_dims, ...), `model.variance_network`, `model.rendering_network`, `model.trainer`
(samples, perturb) and `general.base_exp_dir`; if the file does not exist the constants of
confs/one2345_lod0_val_demo.conf are used.  Weights: `--checkpoint_path`, else -- as the reference does with
`--is_continue` (:137-149) -- the lexicographically last `<base_exp_dir>/checkpoints/ckpt*.pth`, else the seeded synthetic
weights.  Training modes stay with the reference.
"""
import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--conf', type=str, default='./confs/one2345_lod0_val_demo.conf')
    ap.add_argument('--mode', type=str, default='export_mesh')
    ap.add_argument('--threshold', type=float, default=0.0)
    ap.add_argument('--is_continue', default=False, action="store_true")
    ap.add_argument('--is_restore', default=False, action="store_true")
    ap.add_argument('--is_finetune', default=False, action="store_true")
    ap.add_argument('--train_from_scratch', default=False, action="store_true")
    ap.add_argument('--restore_lod0', default=False, action="store_true")
    ap.add_argument('--local_rank', type=int, default=0)
    ap.add_argument('--specific_dataset_name', type=str, default='GSO')
    ap.add_argument('--resolution', type=int, default=360)
    ap.add_argument('--checkpoint_path', type=str, default=None, help='ckpt_*.pth of the reference (default: synthetic weights)')
    args = ap.parse_args(argv)
    if args.mode not in ("export_mesh", "val"):
        raise SystemExit(f"mode={args.mode!r}: only 'export_mesh' and 'val' run on the o2345 path (training stays with the reference)")
    if not torch.cuda.is_available():
        raise SystemExit("needs a CUDA device: the o2345 path has no CPU fallback")
    from o2345 import synthetic as S
    from o2345.checkpoints import latest_checkpoint, load_conf, recon_states
    from o2345.pipeline import build_networks, load_sample
    dev = torch.device("cuda", args.local_rank)
    torch.cuda.set_device(dev)
    exp_dir = args.specific_dataset_name
    note = lambda m: print(m, file=sys.stderr)
    conf = load_conf(args.conf) if os.path.exists(args.conf) else None
    if conf is None:
        note(f"conf {args.conf!r} not found: using the built-in constants of confs/one2345_lod0_val_demo.conf")
    states = S.all_states(0)
    ckpt = args.checkpoint_path
    if ckpt is None and conf is not None and args.is_continue:
        ckpt = latest_checkpoint(conf['general.base_exp_dir'])
        if ckpt is not None:
            note(f"Find checkpoint: {os.path.basename(ckpt)}")
    if ckpt is not None:
        states.update(recon_states(torch.load(ckpt, map_location="cpu"), report=note))
    else:
        note("no checkpoint: seeded synthetic reconstruction weights")
    trainer = build_networks(dev, states=states, base_exp_dir=exp_dir, conf=conf,
                             **({} if conf is not None else {"vol_dim": 96, "perturb": 0.0}))
    sample = load_sample(exp_dir, dev)
    if args.mode == "export_mesh":
        mesh = trainer(sample, mode="export_mesh", resolution=args.resolution)
        print(f"{len(mesh['vertices'])} vertices, {len(mesh['triangles'])} triangles -> {os.path.join(exp_dir, 'mesh.ply')}")
        return mesh
    # perturb_overwrite stays -1: the stratified jitter follows the conf's model.trainer.perturb (1.0 in the demo conf), as in
    # the reference's validate(); white background and alpha_inter_ratio 1.0 as at iter_step 215 000 (:412-418,528-540)
    # (512-ray chunks as in the reference: the host generator's draws -- jitter, then 1024 random points per chunk -- interleave
    # the same way)
    out = trainer(sample, mode="val", background_rgb=1.0, alpha_inter_ratio_lod0=1.0)
    W, H = int(sample['img_wh'][0][0]), int(sample['img_wh'][0][1])
    from PIL import Image
    Image.fromarray((np.clip(out["color"].reshape(H, W, 3), 0, 1) * 255).astype(np.uint8)).save(os.path.join(exp_dir, "val_color.png"))
    np.save(os.path.join(exp_dir, "val_depth.npy"), out["depth"].reshape(H, W))
    np.save(os.path.join(exp_dir, "val_normal.npy"), out["normal"].reshape(H, W, 3))
    print("val outputs written to", exp_dir)
    return out


if __name__ == "__main__":
    main()
