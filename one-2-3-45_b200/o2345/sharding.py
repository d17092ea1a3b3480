"""Image-granularity data parallelism: independent scenes, one process per GPU,
no data-path collective.  Replaces nn.DataParallel in reference exp_runner_generic_blender_val.py:151."""
from __future__ import annotations

import torch
import torch.distributed as dist


def assign_scenes(n_scenes: int, world: int, rank: int):
    """Static round robin: scene i -> rank i mod world."""
    return list(range(rank, n_scenes, world))


def broadcast_module_weights(modules, src=0):
    """The only collective on the path: every parameter and buffer from rank `src` (NCCL over NVLink on GPU)."""
    n = 0
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return n
    for m in modules:
        for t in list(m.parameters()) + list(m.buffers()):
            dist.broadcast(t.data, src=src)
            n += t.numel()
    return n


def max_over_ranks(values, device):
    """MAX-reduce a list of floats (timings) across ranks."""
    t = torch.tensor(values, dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def weights_checksum(modules):
    s = 0.0
    for m in modules:
        for t in list(m.parameters()) + list(m.buffers()):
            s += float(t.detach().double().abs().sum())
    return s
