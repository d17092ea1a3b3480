"""bench.py -- volume-render throughput of the reconstruction hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Metric (BASELINE.json): "volume-render M rays/sec".  One step = GenericTrainer mode='val' on one scene of
BASELINE configs[1]'s reconstruction half: 32 source views of 256x256 -> FeatureNet -> 96^3 cost volume ->
sparse U-Net -> hierarchical ray march of the full 256x256 query image (65 536 rays x (64+64) samples x 32
views), fp32.  `value` keeps the scene resident in HBM; `e2e` runs the same step through the public call
(`trainer(sample, mode='val')`) starting from pinned HOST buffers and ending with the rendered colour /
depth / normal images back on the host.  N > 1: one process per GPU, one independent scene per rank
(weak scaling, no data-path collective; NCCL only broadcasts the weights once and reduces the timing).

NOTE: the other half of BASELINE's metric string (sec/mesh end to end) needs the Zero123 DDIM stage
(SURVEY.md rows A1-A9), which is not built yet; the reconstruction-only mesh time is reported as the
informational key "export_mesh_s" and is NOT a sec/mesh claim.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "one-2-3-45_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

H = W = 256
N_VIEWS = 32
VOL = 96
N_RAYS = H * W
N_S, N_I = 64, 64
CHUNK = 8192            # rays per render() call (the reference uses 512; results are chunk-invariant)
MESH_RES = 256
CONFIG = {"workload": "configs[1] reconstruction half: 32 views 256x256 -> 96^3 volume -> render 65536 rays x (64+64) samples",
          "views": N_VIEWS, "vol_dim": VOL, "rays": N_RAYS, "samples": [N_S, N_I], "chunk_rays": CHUNK,
          "l2": "inputs larger than L2 (feature maps 470 MB + channel-last maps 503 MB + volume 57 MB per step)",
          "parallelism": "one scene per GPU"}
# algorithmic work of SURVEY.md section 8(d)
FLOP_SDF_FWD = 2 * 41856.0
FLOP_SDF_BWD = 2 * (128 * 144 + 128 * 39)
RAY_FLOP = 211e6
RAY_GATHER_BYTES = 4.0e6


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_sustained": p.get("bf16_tflops_sustained"),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def build_scene(dev, seed):
    from o2345 import synthetic as S
    from o2345.pipeline import _sample_from
    cams = S.scene_cameras(S.pose_json(60.0), n_src=N_VIEWS, img_wh=(W, H))
    imgs = S.images(N_VIEWS + 1, H, W, seed=seed)
    return _sample_from(cams, imgs, dev, H, W, pin=True)


def ev_time(fn, stream=None):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b), out


def run_gpu(args):
    import torch.distributed as dist
    from o2345 import _lib, synthetic as S
    from o2345.pipeline import build_networks
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from o2345 import sharding
    tr = build_networks(dev, vol_dim=VOL, states=S.all_states(0), perturb=0.0)
    # the only collective on the path: weights from rank 0 over NVLink (no-op at N = 1)
    sharding.broadcast_module_weights([tr.pyramid_feature_network_geometry_lod0, tr.sdf_network_lod0,
                                       tr.rendering_network_lod0, tr.variance_network_lod0], src=0)
    sample, host, host_rays = build_scene(dev, seed=1234 + rank)
    tr_val = lambda smp: tr.val_step(smp, perturb_overwrite=0, background_rgb=1.0, alpha_inter_ratio_lod0=1.0, chunk_size=CHUNK)

    def step_resident():
        return tr_val_device(tr, sample)

    def step_e2e():
        smp = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        smp["rays"] = {k: v.to(dev, non_blocking=True) for k, v in host_rays.items()}
        smp["batch_idx"], smp["meta"] = sample["batch_idx"], sample["meta"]
        return tr_val(smp)                      # ends with .cpu() of colour / depth / normal

    for _ in range(args.warmup):
        step_resident()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = ClockSampler(local)
    clocks.start()
    _lib.reset_launches()
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        step_resident()
    t1.record()
    torch.cuda.synchronize()
    launches = _lib.launches()
    ms = t0.elapsed_time(t1)
    clk = clocks.stop()
    # e2e: same step from pinned host buffers, results read back to the host every step
    step_e2e()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - w0
    ms, e2e_ms = sharding.max_over_ranks([ms, e2e_s * 1e3], dev)
    out = None
    if rank == 0:
        pk = peaks()
        rays_per_s = world * args.steps * N_RAYS / (ms * 1e-3)
        e2e_rays = world * args.steps * N_RAYS / (e2e_ms * 1e-3)
        h2d = sum(v.numel() * v.element_size() for v in list(host.values()) + list(host_rays.values()))
        d2h = N_RAYS * (3 + 1 + 3) * 4
        roof, mesh_s = kernel_rooflines(tr, sample, dev, pk)
        out = {"metric": "volume-render M rays/sec", "value": rays_per_s / 1e6, "unit": "M rays/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": CONFIG,
               "clocks": clk, "gpu_launches": launches,
               "e2e": {"value": e2e_rays / 1e6, "unit": "M rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
               "roofline": roof, "export_mesh_s": mesh_s, "peaks": pk["source"]}
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(tr, sample, budget_rays=args.cpu_rays)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


def tr_val_device(tr, sample):
    """mode='val' with device-resident inputs and outputs (no host copies inside the step)."""
    imgs, fmaps, cond, sizeW, sizeH = tr._conditional_features(sample)
    vol, occ = cond['dense_volume_scale0'], cond['valid_mask_volume_scale0']
    near, far = sample['query_near_far'][0, :1], sample['query_near_far'][0, 1:]
    ro = sample['rays']['rays_o'][0].reshape(-1, 3)
    rd = sample['rays']['rays_v'][0].reshape(-1, 3)
    outs = []
    for a, b in zip(ro.split(CHUNK), rd.split(CHUNK)):
        o = tr.sdf_renderer_lod0.render(a, b, near, far, tr.sdf_network_lod0, tr.rendering_network_lod0,
                                        perturb_overwrite=0, background_rgb=1.0, alpha_inter_ratio=1.0, lod=0,
                                        conditional_volume=vol, conditional_valid_mask_volume=occ, feature_maps=fmaps,
                                        color_maps=imgs, w2cs=sample['w2cs'][0], intrinsics=sample['intrinsics'][0],
                                        img_wh=[sizeW, sizeH], query_c2w=sample['query_c2w'], if_render_with_grad=False)
        outs.append((o['color_fine'], o['depth']))
    return outs


def kernel_rooflines(tr, sample, dev, pk):
    """Times the dominant kernels in isolation with CUDA events (same stream torch launches on)."""
    from o2345 import ops
    from o2345.sparse_sdf_network import channel_last_volume
    imgs, fmaps, cond, sizeW, sizeH = tr._conditional_features(sample)
    vol, occ = cond['dense_volume_scale0'], cond['valid_mask_volume_scale0']
    vol_cl = channel_last_volume(vol)
    near, far = sample['query_near_far'][0, :1], sample['query_near_far'][0, 1:]
    ro = ops.cf32(sample['rays']['rays_o'][0].reshape(-1, 3))
    rd = ops.cf32(sample['rays']['rays_v'][0].reshape(-1, 3))
    R = ro.shape[0]
    pack = tr.sdf_network_lod0.sdf_layer.packed()
    # a realistic set of fine samples: run the hierarchical sampling once for the whole image
    o = tr.sdf_renderer_lod0.render(ro, rd, near, far, tr.sdf_network_lod0, tr.rendering_network_lod0,
                                    perturb_overwrite=0, background_rgb=1.0, alpha_inter_ratio=1.0, lod=0,
                                    conditional_volume=vol, conditional_valid_mask_volume=occ, feature_maps=fmaps,
                                    color_maps=imgs, w2cs=sample['w2cs'][0], intrinsics=sample['intrinsics'][0],
                                    img_wh=[sizeW, sizeH], query_c2w=sample['query_c2w'])
    mid = o['mid_z_vals'].contiguous()
    active = (o['inside_sphere'] > 0).to(torch.uint8).reshape(-1).contiguous()
    n_act = int(active.sum())
    src = ops.PointSource.rays(ro, rd, mid)
    res = {}
    for name, fn, flop in (
        ("sdf_query_kernel<grad> (fine pass, 65536x128 samples)",
         lambda: ops.sdf_query(src, vol_cl, pack, active=active, want_grad=True), n_act * (FLOP_SDF_FWD + FLOP_SDF_BWD)),
        ("sdf_query_kernel<fwd> (coarse pass, 65536x64 samples)",
         lambda: ops.sdf_query(ops.PointSource.rays(ro, rd, o['z_vals'][:, ::2].contiguous()), vol_cl, pack), R * 64 * FLOP_SDF_FWD),
    ):
        for _ in range(2):
            fn()
        ts = [ev_time(fn)[0] for _ in range(5)]
        res[name] = {"ms": float(np.mean(ts)), "tflops": flop / (np.mean(ts) * 1e-3) / 1e12}
    views = tr.sdf_renderer_lod0._source_views(fmaps, imgs, sample['w2cs'][0], sample['intrinsics'][0], [sizeW, sizeH])
    qc = ops.cf32(sample['query_c2w'].reshape(-1, 4, 4)[0, :3, 3])
    fn = lambda: ops.render_blend(src, active, vol_cl, occ, views, tr.rendering_network_lod0.packed(), query_center=qc)
    for _ in range(2):
        fn()
    ts = [ev_time(fn)[0] for _ in range(5)]
    nvalid = fn()[1]
    pairs = int(nvalid.sum())
    gather = pairs * 4 * 240 + n_act * (8 * 64 + 8 * 4)       # bytes actually requested (valid views only)
    res["render_blend_kernel (65536x128 samples x 32 views)"] = {
        "ms": float(np.mean(ts)), "gather_gbs": gather / (np.mean(ts) * 1e-3) / 1e9, "valid_pairs": pairs, "active_samples": n_act}
    # whole ray march of one image, against SURVEY.md 8(d)'s contract t_roof = max(FLOP/peak, bytes/HBM)
    ms_img, _ = ev_time(lambda: [tr.sdf_renderer_lod0.render(a, b, near, far, tr.sdf_network_lod0, tr.rendering_network_lod0,
                                                              perturb_overwrite=0, background_rgb=1.0, alpha_inter_ratio=1.0, lod=0,
                                                              conditional_volume=vol, conditional_valid_mask_volume=occ,
                                                              feature_maps=fmaps, color_maps=imgs, w2cs=sample['w2cs'][0],
                                                              intrinsics=sample['intrinsics'][0], img_wh=[sizeW, sizeH],
                                                              query_c2w=sample['query_c2w'])
                                 for a, b in zip(ro.split(CHUNK), rd.split(CHUNK))])
    t_roof = max(RAY_FLOP / (pk["bf16_tflops"] * 1e12), RAY_GATHER_BYTES / (pk["hbm_gbs"] * 1e9)) * R
    dom = max(res, key=lambda k: res[k]["ms"])
    d = res[dom]
    if "tflops" in d:
        roof = {"kernel": dom, "bound": "tensor", "achieved": d["tflops"], "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                "frac": d["tflops"] / pk["bf16_tflops"], "traffic": None,
                "note": "fp32 FFMA kernel measured against the bf16 tensor peak (SURVEY.md 8(d) row B8); fp32 SIMT peak is ~72 TFLOP/s"}
    else:
        roof = {"kernel": dom, "bound": "hbm", "achieved": d["gather_gbs"], "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": d["gather_gbs"] / pk["hbm_gbs"], "traffic": None,
                "note": "achieved = requested gather bytes of the valid (sample, view) pairs / kernel time"}
    roof["kernels"] = res
    roof["raymarch_image_ms"] = ms_img
    roof["raymarch_frac_of_contract"] = (t_roof * 1e3) / ms_img
    # informational: reconstruction-only mesh export at R=256 (device part + host copies)
    tr.base_exp_dir = None
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    tr(sample, mode="export_mesh", resolution=MESH_RES)
    torch.cuda.synchronize()
    return roof, time.perf_counter() - w0


def cpu_baseline(tr, sample, budget_rays=48):
    """Oracle (CPU port of the reference) on a bounded sample: `budget_rays` rays of the same image against the
    same (GPU-built) volume and feature maps, all host threads.  A reported baseline, not the target."""
    from helpers import states_torch
    from oracle import recon_oracle as O
    imgs, fmaps, cond, sizeW, sizeH = tr._conditional_features(sample)
    st = states_torch(0)
    vol, occ = cond['dense_volume_scale0'].cpu(), cond['valid_mask_volume_scale0'].cpu()
    sel = torch.linspace(0, N_RAYS - 1, budget_rays).long()
    ro = sample['rays']['rays_o'][0].reshape(-1, 3).cpu()[sel]
    rd = sample['rays']['rays_v'][0].reshape(-1, 3).cpu()[sel]
    near, far = sample['query_near_far'][0, :1].cpu(), sample['query_near_far'][0, 1:].cpu()
    torch.set_num_threads(host_threads())
    w0 = time.perf_counter()
    O.render_rays(ro, rd, near, far, vol, occ, fmaps.cpu(), imgs.cpu(), sample['w2cs'][0].cpu(), sample['intrinsics'][0].cpu(),
                  sample['query_c2w'].cpu(), st["sdf_network_lod0"], st["rendering_network_lod0"],
                  st["variance_network_lod0"]["variance"], W=W, H=H)
    dt = time.perf_counter() - w0
    return {"value": budget_rays / dt / 1e6, "unit": "M rays/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{budget_rays} rays x (64+64) samples x {N_VIEWS} views of the same image, volume + feature maps prebuilt ({dt:.1f} s)"}


def host_threads():
    """Threads for the CPU arm: every core up to 32 (the torch-CPU port stops scaling, and on a 128-thread
    host gets slower, beyond that; the count actually used is what the JSON reports as `cores`)."""
    return max(1, min(os.cpu_count() or 1, 32))


def run_reference(args):
    """--impl reference: the CPU restatement of the reference (oracle/) on the host cores, same metric."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from helpers import states_torch
    from o2345 import synthetic as S
    from oracle import recon_oracle as O
    torch.set_num_threads(host_threads())
    st = states_torch(0)
    cams = S.scene_cameras(S.pose_json(60.0), n_src=N_VIEWS, img_wh=(W, H))
    imgs = torch.from_numpy(S.images(N_VIEWS + 1, H, W, seed=1234))[1:]
    t = lambda x: torch.from_numpy(np.asarray(x)).float()
    w0 = time.perf_counter()
    fm = O.pyramid_feature_maps(imgs, st["pyramid_feature_network"])
    cv = O.conditional_volume(fm, t(cams["partial_vol_origin"]), t(cams["affine_mats"]), st["sdf_network_lod0"], VOL,
                              2.0 / (VOL - 1), H, W)
    t_vol = time.perf_counter() - w0
    ro_all, rv_all = S.query_rays(cams["query_intrinsic"], cams["query_c2w"], H, W)
    n = args.ref_rays
    times = []
    for i in range(args.warmup + args.steps):
        sel = np.linspace(i, N_RAYS - 1 - i, n).astype(np.int64)
        w0 = time.perf_counter()
        O.render_rays(t(ro_all[sel]), t(rv_all[sel]), t(cams["query_near_far"][:1]), t(cams["query_near_far"][1:]),
                      cv["dense"], cv["occ"], fm, imgs, t(cams["w2cs"]), t(cams["intrinsics"]), t(cams["query_c2w"])[None],
                      st["sdf_network_lod0"], st["rendering_network_lod0"], st["variance_network_lod0"]["variance"], W=W, H=H)
        if i >= args.warmup:
            times.append(time.perf_counter() - w0)
    per_ray = float(np.mean(times)) / n
    step_s = t_vol + per_ray * N_RAYS          # one full step = volume build + all 65536 rays (extrapolated)
    val = N_RAYS / step_s / 1e6
    cores = torch.get_num_threads()
    smp = (f"volume build once ({t_vol:.1f} s, timed) + {n} rays per step extrapolated to 65536 rays "
           f"({per_ray * 1e3:.1f} ms/ray)")
    print(json.dumps({"impl": "reference", "metric": "volume-render M rays/sec", "value": val, "unit": "M rays/s",
                      "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                      "data": "synthetic", "config": CONFIG,
                      "cpu_baseline": {"value": val, "unit": "M rays/s", "cores": cores, "kind": "port", "sample": smp},
                      "e2e": {"value": val, "unit": "M rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="o2345", choices=["o2345", "reference"])
    ap.add_argument("--cpu-rays", type=int, default=48)
    ap.add_argument("--ref-rays", type=int, default=24)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the o2345 path has no CPU fallback")
    run_gpu(args)


if __name__ == "__main__":
    main()
