"""Where one image_to_mesh step goes, phase by phase (each phase bracketed by a synchronize: the sum is slightly above the
pipelined step, the shares are what matters)."""
import os, sys, time, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "one-2-3-45_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import bench
from o2345 import synthetic as S, zero123 as Z, ddim as DD
from o2345.pipeline import build_networks, image_to_mesh
import o2345.pipeline as P
dev = torch.device("cuda:0")
tr = build_networks(dev, vol_dim=bench.VOL, states=S.all_states(0), perturb=0.0)
z123 = Z.build_zero123(dev, seed=0, clip=True).half()
img = bench.input_image(4321)
acc = collections.OrderedDict()
def timed(label, fn):
    def wrap(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize(); acc[label] = acc.get(label, 0.0) + time.perf_counter() - t0
        return r
    return wrap
for _ in range(2):
    image_to_mesh(z123, tr, img, polar_angle=60, resolution=bench.MESH_RES)
z123.get_learned_conditioning = timed("CLIP image embedding", z123.get_learned_conditioning)
z123.encode_first_stage = timed("VAE encode", z123.encode_first_stage)
z123.decode_first_stage = timed("VAE decode", z123.decode_first_stage)
_sample = DD.DDIMSampler.sample
DD.DDIMSampler.sample = timed("DDIM sampling (UNet iterations + updates)", _sample)
_gv = Z.generate_views
Z.generate_views = timed("generate_views total", _gv)
P.sample_from_views = timed("views -> cameras / rays / device batch", P.sample_from_views)
tr_call = tr.__class__.__call__
tr.__class__.__call__ = timed("reconstruction (volume, SDF grid, marching cubes, colours, mesh tail)", tr_call)
N = 3
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(N):
    image_to_mesh(z123, tr, img, polar_angle=60, resolution=bench.MESH_RES)
torch.cuda.synchronize(); tot = (time.perf_counter() - t0) / N
for k, v in acc.items():
    print("%-75s %8.1f ms" % (k, v / N * 1e3))
inner = sum(v for k, v in acc.items() if k not in ("generate_views total",)) / N
gv = acc["generate_views total"] / N
parts = sum(acc[k] for k in ("CLIP image embedding", "VAE encode", "VAE decode", "DDIM sampling (UNet iterations + updates)")) / N
print("%-75s %8.1f ms" % ("generate_views: host glue (noise draws, conditioning, uint8 hand-off, .cpu())", (gv - parts) * 1e3))
print("%-75s %8.1f ms" % ("step total (with the synchronizes)", tot * 1e3))
# the UNet graph alone, in this process on this GPU (box-to-box drift is ~5 %): what the sampler's iterations cost without its glue
unet = z123.model.diffusion_model
for B, n_it in bench.UNET_SCHEDULE:
    x = torch.randn(B, 8, 32, 32, device=dev); t = torch.full((B,), 501, device=dev); ctx = torch.randn(B, 1, 768, device=dev)
    unet(x, t, ctx); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): unet(x, t, ctx)
    e1.record(); torch.cuda.synchronize()
    print("UNet graph at batch %d: %.3f ms per iteration x %d = %.1f ms" % (B, e0.elapsed_time(e1) / 10, n_it, e0.elapsed_time(e1) / 10 * n_it))
vae = z123.first_stage_model
zz = torch.randn(4, 4, 32, 32, device=dev)
vae.decode(zz); torch.cuda.synchronize()
e0.record()
for _ in range(5): vae.decode(zz)
e1.record(); torch.cuda.synchronize()
print("VAE decode(4): %.3f ms" % (e0.elapsed_time(e1) / 5))
