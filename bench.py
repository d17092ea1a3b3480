"""bench.py -- One-2-3-45 hot paths on B200: sec/mesh end to end, and volume-render M rays/sec.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Headline metric (BASELINE.json, configs[1]): "sec/mesh end-to-end (256x256 in)".  One step = one 256x256 input
image -> Zero123 stage 1 + stage 2 (10 DDIM sampler calls = 2x76 + 8x49 = 544 UNet iterations at the CFG batch
of 8, fp16 tensor-core GEMMs with fp32 accumulate; 10 CLIP ViT-L/14 image embeddings, 10 VAE encodes, 40 VAE decodes)
-> 32 views -> FeatureNet ->
96^3 cost volume -> sparse U-Net -> 256^3 SDF grid -> marching cubes -> vertex colours -> mesh arrays on the host
(`o2345.pipeline.image_to_mesh`).  Not inside the step (out of scope): SAM / rembg
preprocessing and the LoFTR elevation search (polar angle 60).
`value` is timed on the device with CUDA events; `e2e` is the wall clock of the same public call starting from a
pinned host image and ending with the mesh on the host (the pipeline itself moves the generated views through the
host as uint8, as the reference's PNG hand-off does).  The second BASELINE metric, "volume-render M rays/sec", is
reported under "rays" with its own roofline (GenericTrainer mode='val' on 65 536 rays x (64+64) samples x 32 views).
N > 1: one process per GPU, one independent image per rank (weak scaling, no data-path collective).
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
CHUNK = 8192
MESH_RES = 256
UNET_ITERS = 2 * 76 + 8 * 49
CONFIG = {"workload": "configs[1]: single 256x256 image -> mesh: Zero123 75/50-step DDIM fp16 (544 UNet iterations at batch 8) "
                      "+ 96^3 cost volume + 256^3 SDF grid + marching cubes, 1 image per GPU",
          "views": N_VIEWS, "vol_dim": VOL, "mesh_resolution": MESH_RES, "ddim_steps": [75, 50], "cfg_scale": 3.0,
          "l2": "inputs larger than L2 (1.72 GB fp16 UNet weights stream every iteration; 470 MB feature maps)",
          "not_in_step": "SAM/rembg, LoFTR elevation search (polar angle 60)",
          "parallelism": "one image per GPU"}
# algorithmic work, SURVEY.md section 8(d)
UNET_FLOP_PER_SAMPLE = 176.3e9
FLOP_SDF_FWD = 2 * 41856.0
FLOP_SDF_BWD = 2 * (128 * 144 + 128 * 39)
RAY_FLOP, RAY_GATHER_BYTES = 211e6, 4.0e6
PUBLISHED_SEC_PER_MESH = 40.0   # BASELINE.md section 1 (reference README.md:154, A6000, whole run.py)


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
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
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
        ok = [r for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        sm = [float(r[1]) for r in ok]
        mx = [float(r[2]) for r in ok]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in ok for i in range(4) if r[4 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def input_image(seed):
    from o2345 import synthetic as S
    return (S.images(1, H, W, seed=seed)[0].transpose(1, 2, 0) * 255.0).astype(np.uint8)


def ev_time(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b), out


def run_gpu(args):
    import torch.distributed as dist
    from o2345 import _lib, sharding, synthetic as S
    from o2345.pipeline import build_networks, image_to_mesh
    from o2345.zero123 import build_zero123
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep NCCL's version banner / warnings off stdout: ONE JSON line is the contract
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    tr = build_networks(dev, vol_dim=VOL, states=S.all_states(0), perturb=0.0)
    z123 = build_zero123(dev, seed=0, clip=True).half()   # `--half_precision`: fp16-rounded schedule buffers; CLIP tower attached
    # the only collective on the path: weights from rank 0 over NVLink (no-op at N = 1)
    sharding.broadcast_module_weights([tr.pyramid_feature_network_geometry_lod0, tr.sdf_network_lod0,
                                       tr.rendering_network_lod0, tr.variance_network_lod0, z123], src=0)
    img_host = torch.from_numpy(input_image(4321 + rank)).pin_memory()
    step = lambda: image_to_mesh(z123, tr, img_host.numpy(), polar_angle=60, resolution=MESH_RES)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = ClockSampler(local)
    clocks.start()
    _lib.reset_launches()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    t0.record()
    for _ in range(args.steps):
        mesh = step()
    t1.record()
    torch.cuda.synchronize()
    wall_s = time.perf_counter() - w0
    launches = _lib.launches()
    clk = clocks.stop()
    ms, e2e_ms = sharding.max_over_ranks([t0.elapsed_time(t1), wall_s * 1e3], dev)
    if rank == 0:
        pk = peaks()
        sec_per_mesh = ms * 1e-3 / (args.steps * world)
        out = {"metric": "sec/mesh end-to-end (256x256 in)", "value": sec_per_mesh, "unit": "s/mesh", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": False,
               "scaling": "weak", "vs_baseline": sec_per_mesh / PUBLISHED_SEC_PER_MESH, "dtype": "f16", "data": "synthetic",
               "config": CONFIG, "baseline_note": "BASELINE.md section 1: 40 s per image for run.py --half_precision on an A6000 "
                                                  "(reference README.md:154), which also covers SAM, the LoFTR elevation "
                                                  "search, model loading and a second process start -- not in this step",
               "clocks": clk, "gpu_launches": launches,
               "e2e": {"value": e2e_ms * 1e-3 / (args.steps * world), "unit": "s/mesh", "h2d_bytes_per_step": int(img_host.numel()),
                       "d2h_bytes_per_step": int(mesh["vertices"].nbytes + mesh["triangles"].nbytes + mesh["colors"].nbytes)},
               "mesh": {"vertices": int(len(mesh["vertices"])), "triangles": int(len(mesh["triangles"]))},
               "peaks": pk["source"]}
        out.update(stage_breakdown(z123, tr, dev, pk))
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_reference()
        emit(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def stage_breakdown(z123, tr, dev, pk):
    """Per-stage device times, the tensor-core roofline of the dominant kernel (the tcgen05 GEMM inside the UNet)
    and the volume-rendering throughput with its own roofline."""
    from o2345 import ops_a
    from o2345.pipeline import synthetic_sample
    unet, vae = z123.model.diffusion_model, z123.first_stage_model
    x = torch.randn(8, 8, 32, 32, device=dev)
    t = torch.full((8,), 501, device=dev)
    ctx = torch.randn(8, 1, 768, device=dev)
    unet(x, t, ctx)
    ms_unet = float(np.mean([ev_time(lambda: unet(x, t, ctx))[0] for _ in range(10)]))
    # Device time of the tensor-core kernel inside one UNet pass: every GEMM / implicit-conv call of an eager pass is
    # recorded (operands kept alive) and replayed back to back inside ONE CUDA graph, timed with events around the
    # replay -- the kernel's launches exactly as the captured UNet graph issues them, without the glue kernels between.
    rec = []
    real = {n: getattr(ops_a, n) for n in ("gemm", "bgemm", "conv3x3")}

    def spy(name):
        def wrap(*a, **k):
            rec.append((name, a, k))
            return real[name](*a, **k)
        return wrap
    for n in real:
        setattr(ops_a, n, spy(n))
    unet.use_cuda_graph = False
    try:
        unet(x, t, ctx)
        torch.cuda.synchronize()
    finally:
        for n in real:
            setattr(ops_a, n, real[n])
        unet.use_cuda_graph = True
    flops_counted = 0.0
    for name, a, k in rec:
        if name == "gemm":
            flops_counted += 2.0 * a[0].shape[0] * a[0].shape[1] * a[1].shape[0]
        elif name == "conv3x3":
            flops_counted += 2.0 * a[1] * a[2] * a[3] * 9 * a[4] * a[5].shape[0]
        else:
            flops_counted += 2.0 * a[3] * a[4] * a[8] * a[9] * a[10]
    side = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        for name, a, k in rec:
            real[name](*a, **k)
        torch.cuda.synchronize()
        with torch.cuda.graph(graph, stream=side):
            for name, a, k in rec:
                real[name](*a, **k)
    graph.replay()
    torch.cuda.synchronize()
    ms_gemm = float(np.median([ev_time(graph.replay)[0] for _ in range(10)]))
    n_gemm = len(rec)
    del graph
    z = torch.randn(4, 4, 32, 32, device=dev)
    vae.decode(z)
    ms_dec = float(np.mean([ev_time(lambda: vae.decode(z))[0] for _ in range(3)]))
    flops = 8 * UNET_FLOP_PER_SAMPLE
    tf = flops_counted / (ms_gemm * 1e-3) / 1e12
    roofline = {"kernel": "gemm2_f16_tc_kernel (tcgen05.mma.cta_group::2 kind::f16; all %d GEMM / implicit-conv launches of one "
                          "UNet iteration at batch 8)" % n_gemm,
                "bound": "tensor", "achieved": tf, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": tf / pk["bf16_tflops"],
                "traffic": None, "flops_per_step": flops_counted, "gemm_ms_per_unet_iteration": ms_gemm,
                "note": "algorithmic FLOPs = sum of 2 M N K over the recorded launches (%.1f GFLOP; SURVEY.md 8(d) row A2 quotes "
                        "%.1f GFLOP for the same pass including attention) / CUDA-event time of those launches replayed back to back "
                        "in one CUDA graph (split-K finalize kernels included)" % (flops_counted / 1e9, flops / 1e9)}
    # ---- reconstruction stages + volume rendering
    sample = synthetic_sample(dev, n_views=N_VIEWS, H=H, W=W)
    tr._conditional_features(sample)
    ms_front, (imgs, fmaps, cond, sizeW, sizeH) = ev_time(lambda: tr._conditional_features(sample))
    tr.base_exp_dir = None
    tr(sample, mode="export_mesh", resolution=MESH_RES)
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    tr(sample, mode="export_mesh", resolution=MESH_RES)
    torch.cuda.synchronize()
    s_mesh = time.perf_counter() - w0
    rays = render_throughput(tr, sample, imgs, fmaps, cond, sizeW, sizeH, dev, pk)
    stages = {"unet_iteration_ms": ms_unet, "unet_total_s": ms_unet * UNET_ITERS * 1e-3, "vae_decode4_ms": ms_dec,
              "volume_build_ms": ms_front, "export_mesh_s": s_mesh}
    return {"roofline": roofline, "stages": stages, "rays": rays}


def render_throughput(tr, sample, imgs, fmaps, cond, sizeW, sizeH, dev, pk):
    from o2345 import ops
    from o2345.sparse_sdf_network import channel_last_volume
    vol, occ = cond['dense_volume_scale0'], cond['valid_mask_volume_scale0']
    near, far = sample['query_near_far'][0, :1], sample['query_near_far'][0, 1:]
    ro = ops.cf32(sample['rays']['rays_o'][0].reshape(-1, 3))
    rd = ops.cf32(sample['rays']['rays_v'][0].reshape(-1, 3))

    def image():
        outs = []
        for a, b in zip(ro.split(CHUNK), rd.split(CHUNK)):
            o = tr.sdf_renderer_lod0.render(a, b, near, far, tr.sdf_network_lod0, tr.rendering_network_lod0,
                                            perturb_overwrite=0, background_rgb=1.0, alpha_inter_ratio=1.0, lod=0,
                                            conditional_volume=vol, conditional_valid_mask_volume=occ, feature_maps=fmaps,
                                            color_maps=imgs, w2cs=sample['w2cs'][0], intrinsics=sample['intrinsics'][0],
                                            img_wh=[sizeW, sizeH], query_c2w=sample['query_c2w'])
            outs.append(o)
        return outs
    image()
    ms_img, outs = ev_time(image)
    o = outs[0]
    # dominant kernels of the ray march, timed alone on the first chunk
    vol_cl = channel_last_volume(vol)
    pack = tr.sdf_network_lod0.sdf_layer.packed()
    mid = o['mid_z_vals'].contiguous()
    active = (o['inside_sphere'] > 0).to(torch.uint8).reshape(-1).contiguous()
    n_act = int(active.sum())
    src = ops.PointSource.rays(ro[:CHUNK], rd[:CHUNK], mid)
    f_sdf = lambda: ops.sdf_query(src, vol_cl, pack, active=active, want_grad=True)
    f_sdf()
    ms_sdf = float(np.mean([ev_time(f_sdf)[0] for _ in range(3)]))
    views = tr.sdf_renderer_lod0._source_views(fmaps, imgs, sample['w2cs'][0], sample['intrinsics'][0], [sizeW, sizeH])
    qc = ops.cf32(sample['query_c2w'].reshape(-1, 4, 4)[0, :3, 3])
    prec = tr.sdf_renderer_lod0.blend_precision
    f_bl = lambda: ops.render_blend(src, active, vol_cl, occ, views, tr.rendering_network_lod0.packed(), query_center=qc,
                                    precision=prec)
    f_bl()
    ms_bl = float(np.mean([ev_time(f_bl)[0] for _ in range(3)]))
    pairs = int(f_bl()[1].sum())
    f_bl32 = lambda: ops.render_blend(src, active, vol_cl, occ, views, tr.rendering_network_lod0.packed(), query_center=qc,
                                      precision=0)
    rgb32 = f_bl32()[0]
    ms_bl32 = float(np.mean([ev_time(f_bl32)[0] for _ in range(2)]))
    drift = float((f_bl()[0] - rgb32).abs().max())
    t_roof = max(RAY_FLOP / (pk["bf16_tflops"] * 1e12), RAY_GATHER_BYTES / (pk["hbm_gbs"] * 1e9)) * N_RAYS
    return {"metric": "volume-render M rays/sec", "value": N_RAYS / (ms_img * 1e-3) / 1e6, "unit": "M rays/s",
            "workload": "65536 rays x (64+64) samples x 32 views, volume + feature maps resident; SDF MLP split-fp16 tensor cores (fp32-grade), view-blending "
                        "MLPs on tensor cores (fp16 operands, fp32 accumulate / statistics)",
            "image_ms": ms_img, "frac_of_survey_contract": t_roof * 1e3 / ms_img,
            "kernels_first_chunk": {
                "sdf_query_kernel<grad>": {"ms": ms_sdf, "tflops": n_act * (FLOP_SDF_FWD + FLOP_SDF_BWD) / (ms_sdf * 1e-3) / 1e12,
                                           "active_samples": n_act},
                "render_blend_tc_kernel": {"ms": ms_bl, "valid_pairs": pairs, "ms_fp32_kernel": ms_bl32,
                                           "max_colour_drift_vs_fp32_kernel": drift,
                                        "gather_gbs": (pairs * 960 + n_act * 544) / (ms_bl * 1e-3) / 1e9}}}


def host_threads():
    """Threads for the CPU arm: every core up to 32 (the torch-CPU port stops scaling beyond that; the count actually
    used is what the JSON reports as `cores`)."""
    return max(1, min(os.cpu_count() or 1, 32))


def cpu_reference():
    """The reference's own algorithm on the host cores (oracle/ port, fp32): a bounded sample of every stage of one
    mesh, extrapolated to the full workload -- 1 UNet iteration at batch 8 (x544), 1 VAE decode of one latent (x40),
    1 VAE encode (x10), the full 96^3 volume build, and a 48^3 SDF grid (x (256/48)^3)."""
    from helpers import states_torch
    from o2345 import synthetic as S
    from oracle import ldm_oracle as LO, recon_oracle as O, vae_oracle as VO
    torch.set_num_threads(host_threads())
    t = lambda x: torch.from_numpy(np.asarray(x)).float()
    w0 = time.perf_counter()
    sd_u = {k: torch.from_numpy(v) for k, v in S.unet_state(0).items()}
    sd_v = {k: torch.from_numpy(v) for k, v in S.vae_state(10).items()}
    t_setup = time.perf_counter() - w0
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        w0 = time.perf_counter()
        LO.unet_forward(sd_u, torch.randn(8, 8, 32, 32, generator=g), torch.full((8,), 501), torch.randn(8, 1, 768, generator=g))
        t_unet = time.perf_counter() - w0
        w0 = time.perf_counter()
        VO.decode(sd_v, torch.randn(1, 4, 32, 32, generator=g))
        t_dec = time.perf_counter() - w0
        w0 = time.perf_counter()
        VO.encode_moments(sd_v, torch.rand(1, 3, 256, 256, generator=g) * 2 - 1)
        t_enc = time.perf_counter() - w0
        from oracle import clip_oracle as CO
        sd_c = {k: torch.from_numpy(v) for k, v in S.clip_state(20).items()}
        w0 = time.perf_counter()
        CO.embed(sd_c, torch.rand(1, 3, 256, 256, generator=g) * 2 - 1)
        t_clip = time.perf_counter() - w0
        st = states_torch(0)
        cams = S.scene_cameras(S.pose_json(60.0), n_src=N_VIEWS, img_wh=(W, H))
        imgs = torch.from_numpy(S.images(N_VIEWS + 1, H, W, seed=1234))[1:]
        w0 = time.perf_counter()
        fm = O.pyramid_feature_maps(imgs, st["pyramid_feature_network"])
        cv = O.conditional_volume(fm, t(cams["partial_vol_origin"]), t(cams["affine_mats"]), st["sdf_network_lod0"], VOL,
                                  2.0 / (VOL - 1), H, W)
        t_vol = time.perf_counter() - w0
        w0 = time.perf_counter()
        O.sdf_grid(cv["dense"], st["sdf_network_lod0"], 48)
        t_grid = (time.perf_counter() - w0) * (MESH_RES / 48.0) ** 3
    total = UNET_ITERS * t_unet + 40 * t_dec + 10 * t_enc + 10 * t_clip + t_vol + t_grid
    return {"value": total, "unit": "s/mesh", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"1 UNet iteration at batch 8 ({t_unet:.2f} s, x{UNET_ITERS}) + 1 VAE decode ({t_dec:.2f} s, x40) + 1 VAE encode "
                      f"({t_enc:.2f} s, x10) + 1 CLIP image embedding ({t_clip:.2f} s, x10) + full 96^3 volume build ({t_vol:.1f} s) + 48^3 SDF grid scaled to 256^3 ({t_grid:.0f} s); "
                      f"marching cubes / vertex colours not included; weights generated in {t_setup:.0f} s (untimed)"}


def run_reference(args):
    """--impl reference: the CPU port of the reference (oracle/) on the host cores, same metric and config."""
    if int(os.environ.get("RANK", 0)) != 0:
        return
    cb = cpu_reference()
    emit(json.dumps({"impl": "reference", "metric": "sec/mesh end-to-end (256x256 in)", "value": cb["value"], "unit": "s/mesh",
                      "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["value"] * 1e3,
                      "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": CONFIG, "cpu_baseline": cb,
                      "e2e": {"value": cb["value"], "unit": "s/mesh", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


_JSON_OUT = None


def emit(line):
    """The ONE JSON line goes to the process's original stdout; everything else printed to fd 1 (NCCL's version banner,
    library chatter) has been re-routed to stderr by main()."""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(line + "\n")
    out.flush()


def main():
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="o2345", choices=["o2345", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the o2345 path has no CPU fallback")
    run_gpu(args)


if __name__ == "__main__":
    main()
