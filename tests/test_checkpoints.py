"""CPU: the host-side boundary logic for real checkpoints and confs (o2345/checkpoints.py, o2345/mesh_io.py):
HOCON subset, latest-checkpoint discovery, per-network state dicts, EMA-shadow mapping, CLIP key loading, .obj / .glb."""
import json
import os
import struct

import numpy as np
import pytest
import torch

CONF = """
# comment line
general {
  base_exp_dir = %s # trailing comment
  recording = [
    ./,
    ./data
    ./models
  ]
}
dataset { test_img_wh = [256, 256]
  clean_image = True }
train { learning_rate = 2e-4, anneal_end = 25000
  use_white_bkgd = True }
model {
  num_lods = 1
  sdf_network_lod0 {
    lod = 0,
    ch_in = 56,  # the channel num of fused pyramid features
    voxel_size = 0.02105263,
    vol_dims = [96, 96, 96],
    hidden_dim = 128,
    cost_type = variance_mean
    d_pyramid_feature_compress = 16,
    regnet_d_out = 16,
    num_sdf_layers = 4,
    multires = 6
  }
  variance_network { init_val = 0.2 }
  rendering_network {
    in_geometry_feat_ch = 16
    in_rendering_feat_ch = 56
    anti_alias_pooling = True
  }
  trainer { n_samples_lod0 = 64, n_importance_lod0 = 64, n_samples_lod1 = 64, n_importance_lod1 = 64
    n_outside = 0  // 128 if render_outside_uniform_sampling
    perturb = 1.0
    alpha_type = div }
}
"""


def test_conf_subset_of_hocon(tmp_path):
    from o2345.checkpoints import parse_conf
    c = parse_conf(CONF % "exp/lod0")
    assert c['general.base_exp_dir'] == "exp/lod0"
    assert c['general.recording'] == ["./", "./data", "./models"]
    assert c['dataset.test_img_wh'] == [256, 256] and c.get_bool('dataset.clean_image') is True
    assert c.get_float('train.learning_rate') == 2e-4 and c.get_float('train.anneal_end', default=0) == 25000
    assert c.get_int('model.num_lods') == 1
    assert c.get_int('model.h_patch_size', default=3) == 3              # the renderer's optional key (sparse_neus_renderer.py:62)
    with pytest.raises(KeyError):
        c.get_int('model.h_patch_size')
    net = c['model.sdf_network_lod0']
    assert net == {"lod": 0, "ch_in": 56, "voxel_size": 0.02105263, "vol_dims": [96, 96, 96], "hidden_dim": 128,
                   "cost_type": "variance_mean", "d_pyramid_feature_compress": 16, "regnet_d_out": 16, "num_sdf_layers": 4,
                   "multires": 6}
    assert c['model.trainer']['perturb'] == 1.0 and c['model.trainer']['alpha_type'] == "div" and c['model.trainer']['n_outside'] == 0
    c['general.base_exp_dir'] = "elsewhere"                              # the runner overwrites it (exp_runner...val.py:51)
    assert c['general']['base_exp_dir'] == "elsewhere"


def test_latest_checkpoint_and_recon_round_trip(tmp_path):
    """A synthetic ckpt_000123.pth written the way save_checkpoint does is found, loaded and reaches the modules."""
    from o2345 import synthetic as S
    from o2345.checkpoints import latest_checkpoint, parse_conf, recon_states
    from o2345.pipeline import build_networks
    base = tmp_path / "exp" / "lod0"
    (base / "checkpoints").mkdir(parents=True)
    assert latest_checkpoint(str(base)) is None
    st = S.all_states(3)
    ck = {"optimizer": {"state": {}}, "iter_step": 123, "val_step": 0,
          "sdf_network_lod0": {k: torch.as_tensor(np.asarray(v)) for k, v in st["sdf_network_lod0"].items()},
          "sdf_network_lod1": None,
          "rendering_network_lod0": {k: torch.as_tensor(np.asarray(v)) for k, v in st["rendering_network_lod0"].items()},
          "variance_network_lod0": {"variance": torch.tensor(0.123)},
          "pyramid_feature_network": {k: torch.as_tensor(np.asarray(v)) for k, v in st["pyramid_feature_network"].items()}}
    torch.save(ck, base / "checkpoints" / "ckpt_000123.pth")
    torch.save({}, base / "checkpoints" / "ckpt_000045.pth")
    (base / "checkpoints" / "notes.txt").write_text("x")
    (base / "checkpoints" / "ckpt_999999.tmp").write_text("x")
    found = latest_checkpoint(str(base))
    assert os.path.basename(found) == "ckpt_000123.pth"
    said = []
    states = S.all_states(0)
    states.update(recon_states(torch.load(found, map_location="cpu"), report=said.append))
    assert said == []
    conf = parse_conf(CONF % str(base))
    tr = build_networks("cpu", states=states, conf=conf, base_exp_dir=str(tmp_path))
    assert abs(float(tr.variance_network_lod0.variance) - 0.123) < 1e-7
    assert tr.sdf_network_lod0.voxel_size == 0.02105263                   # the conf's 8-digit constant, not 2/95
    assert tr.sdf_renderer_lod0.perturb == 1.0 and tr.sdf_renderer_lod0.n_importance == 64
    w = dict(tr.rendering_network_lod0.named_parameters())
    k0 = next(iter(st["rendering_network_lod0"]))
    assert torch.equal(w[k0], torch.as_tensor(np.asarray(st["rendering_network_lod0"][k0])))
    # a file that lacks a network is reported the way the reference prints it, and the initialisation is kept
    said = []
    part = recon_states({"sdf_network_lod0": ck["sdf_network_lod0"]}, report=said.append)
    assert set(part) == {"sdf_network_lod0"} and "rendering_network_lod0 load fails" in said


TINY_UNET = dict(model_channels=32, channel_mult=(1, 2), attention_resolutions=(1, 2), num_heads=4, context_dim=768)
TINY_VAE = dict(ddconfig=dict(ch=32, ch_mult=(1, 2), num_res_blocks=1))


def _lightning_checkpoint(with_ema=True, drop=None):
    """state_dict with the key layout of zero123-xl.ckpt: model.*, model_ema.* (dots stripped), first_stage_model.*,
    cc_projection.*, cond_stage_model.model.* (vision tower + a few text-side leftovers), schedule buffers."""
    from o2345.zero123 import LatentDiffusion
    g = torch.Generator().manual_seed(0)
    m = LatentDiffusion(unet_config=TINY_UNET, first_stage_config=TINY_VAE)
    sd = {}
    for k, v in m.state_dict().items():
        sd[k] = torch.randn(v.shape, generator=g) if v.dtype.is_floating_point else v.clone()
    if with_ema:
        for n, p in m.model.named_parameters():
            sd["model_ema." + n.replace(".", "")] = sd["model." + n] + 1.0          # EMA differs from the raw weights
        sd["model_ema.decay"], sd["model_ema.num_updates"] = torch.tensor(0.9999), torch.tensor(7)
    sd["sqrt_alphas_cumprod"], sd["logvar"] = torch.zeros(1000), torch.zeros(1000)
    sd["cond_stage_model.model.logit_scale"] = torch.tensor(1.0)                      # text-side leftovers of clip.load
    sd["cond_stage_model.model.token_embedding.weight"] = torch.zeros(4, 4)
    if drop:
        sd = {k: v for k, v in sd.items() if not k.startswith(drop)}
    return sd, m


def test_zero123_checkpoint_loads_the_ema_shadow(tmp_path):
    from o2345.zero123 import load_zero123_checkpoint
    sd, _ = _lightning_checkpoint()
    path = tmp_path / "zero123-tiny.ckpt"
    torch.save({"state_dict": sd, "global_step": 1}, path)
    said = []
    m = load_zero123_checkpoint(str(path), "cpu", unet_config=TINY_UNET, first_stage_config=TINY_VAE, clip=False, report=said.append)
    name = "diffusion_model.input_blocks.1.0.in_layers.2.weight"
    assert torch.equal(dict(m.model.named_parameters())[name], sd["model_ema." + name.replace(".", "")])
    assert torch.equal(m.first_stage_model.state_dict()["decoder.conv_in.weight"], sd["first_stage_model.decoder.conv_in.weight"])
    assert any("EMA" in s for s in said)
    # --no_ema: the raw weights
    m2 = load_zero123_checkpoint(sd, "cpu", use_ema=False, unet_config=TINY_UNET, first_stage_config=TINY_VAE, clip=False, report=None)
    assert torch.equal(dict(m2.model.named_parameters())[name], sd["model." + name])
    # no shadow in the file: the stored weights, and the report says so
    sd3, _ = _lightning_checkpoint(with_ema=False)
    said = []
    m3 = load_zero123_checkpoint(sd3, "cpu", unet_config=TINY_UNET, first_stage_config=TINY_VAE, clip=False, report=said.append)
    assert torch.equal(dict(m3.model.named_parameters())[name], sd3["model." + name]) and any("no EMA" in s for s in said)


def test_zero123_checkpoint_refuses_incomplete_files():
    from o2345.zero123 import load_zero123_checkpoint
    sd, _ = _lightning_checkpoint()
    del sd["model_ema." + "diffusion_model.out.2.weight".replace(".", "")]        # a partial shadow
    with pytest.raises(KeyError, match="EMA shadow"):
        load_zero123_checkpoint(sd, "cpu", unet_config=TINY_UNET, first_stage_config=TINY_VAE, clip=False, report=None)
    sd, _ = _lightning_checkpoint(drop="first_stage_model.decoder")
    with pytest.raises(KeyError, match="sampling path needs"):
        load_zero123_checkpoint(sd, "cpu", unet_config=TINY_UNET, first_stage_config=TINY_VAE, clip=False, report=None)
    sd, _ = _lightning_checkpoint()                                                # no CLIP vision tower in the file
    with pytest.raises(KeyError, match="cond_stage_model.model.visual"):
        load_zero123_checkpoint(sd, "cpu", unet_config=TINY_UNET, first_stage_config=TINY_VAE, clip=True, report=None)


def test_mesh_tail_merge_and_formats(tmp_path):
    from o2345.mesh_io import convert_mesh_format, merge_vertices, read_ply, to_viewer_frame, write_ply
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 0, 0], [1, 1, 0], [5, 5, 5]], np.float64)   # 3 duplicates 1; 5 unreferenced
    f = np.array([[0, 1, 2], [3, 4, 2]])
    c = (np.arange(18).reshape(6, 3) * 10).astype(np.uint8)
    v2, f2, c2 = merge_vertices(v, f, c)
    assert len(v2) == 4 and f2.max() == 3
    assert np.array_equal(v2[f2], v[f])                                   # same triangles in space
    assert np.array_equal(c2[1], c[1])                                    # the first occurrence keeps its colour
    write_ply(str(tmp_path / "mesh.ply"), v2, f2, c2)
    rv, rf, rc = read_ply(str(tmp_path / "mesh.ply"))
    assert np.allclose(rv, v2) and np.array_equal(rf, f2) and np.array_equal(rc[:, :3], c2)
    # reference utils/utils.py:35-41: Rx(+90), Rz(180), x -> -x, reversed winding == (x, y, z) -> (x, z, y)
    tv, tf = to_viewer_frame(np.array([[1.0, 2.0, 3.0]]), np.array([[0, 1, 2]]))
    assert np.allclose(tv, [[1.0, 3.0, 2.0]]) and tf.tolist() == [[2, 1, 0]]
    obj = convert_mesh_format(str(tmp_path), ".obj")
    lines = open(obj).read().splitlines()
    vs = [l.split() for l in lines if l.startswith("v ")]
    assert len(vs) == 4 and len(vs[0]) == 7 and abs(float(vs[1][4]) - c2[1][0] / 255.0) < 1e-6
    assert [l for l in lines if l.startswith("f ")][0] == "f %d %d %d" % tuple(int(i) + 1 for i in f2[0][::-1])
    glb = convert_mesh_format(str(tmp_path), ".glb")
    raw = open(glb, "rb").read()
    magic, version, total = struct.unpack("<III", raw[:12])
    assert magic == 0x46546C67 and version == 2 and total == len(raw)
    jlen, jtype = struct.unpack("<II", raw[12:20])
    doc = json.loads(raw[20:20 + jlen])
    assert jtype == 0x4E4F534A and doc["accessors"][0]["count"] == 4 and doc["accessors"][2]["count"] == 6
