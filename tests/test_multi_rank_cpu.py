"""CPU, gloo, world_size 2: the N > 1 plumbing of bench.py (scene sharding, weight broadcast, MAX timing)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, os.path.join(ROOT, "one-2-3-45_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from o2345 import sharding, synthetic as S
    from o2345.pipeline import build_networks
    tr = build_networks("cpu", vol_dim=24, states=S.all_states(rank))      # ranks start with DIFFERENT weights
    mods = [tr.pyramid_feature_network_geometry_lod0, tr.sdf_network_lod0, tr.rendering_network_lod0, tr.variance_network_lod0]
    before = sharding.weights_checksum(mods)
    n = sharding.broadcast_module_weights(mods, src=0)
    after = sharding.weights_checksum(mods)
    tmax = sharding.max_over_ranks([10.0 + rank, 5.0 - rank], "cpu")
    out[rank] = (before, after, n, tmax, sharding.assign_scenes(7, world, rank))
    dist.destroy_process_group()


def test_two_ranks_share_weights_and_split_scenes():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, 29517, out), nprocs=world, join=True)
    (b0, a0, n0, t0, s0), (b1, a1, n1, t1, s1) = out[0], out[1]
    assert b0 != b1 and a0 == a1 == b0 and n0 == n1 > 100000
    assert t0 == t1 == [11.0, 5.0]
    assert s0 == [0, 2, 4, 6] and s1 == [1, 3, 5]
