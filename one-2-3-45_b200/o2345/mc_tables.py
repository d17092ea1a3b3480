"""Marching-cubes case tables, generated (not transcribed).

PyMCubes (the reference's mesh extractor, reference sparse_neus_renderer.py:932) is not
vendored and its lookup tables are not available offline, so the 256-case triangle table
is derived here from first principles:

* corner ``i`` of a cell sits at offset ``(i&1, (i>>1)&1, (i>>2)&1)`` in (x, y, z);
* edge ``e = 4*axis + j`` runs along ``axis`` from the corner whose other two coordinates
  are the bits of ``j`` (lower axis = bit 0);
* on every cube face the crossed edges are joined pairwise; on an ambiguous face (two
  diagonal inside corners) each *inside* corner is cut off on its own, a rule that depends
  only on the face's four corner signs, so neighbouring cells always agree and the surface
  is watertight;
* the closed loops are fan-triangulated and wound so that normals point from the inside
  (``u > iso``) to the outside.

The set of mesh vertices (one per sign-changing lattice edge) does not depend on the
table at all; only the triangle connectivity inside ambiguous cells can differ from
PyMCubes' transcription of the classic table.
"""
from __future__ import annotations

import functools

import numpy as np

CORNERS = np.array([[(i >> 0) & 1, (i >> 1) & 1, (i >> 2) & 1] for i in range(8)], np.int32)


def _edge_endpoints():
    ends = np.zeros((12, 2), np.int32)
    for axis in range(3):
        others = [a for a in range(3) if a != axis]
        for j in range(4):
            c = [0, 0, 0]
            c[others[0]] = j & 1
            c[others[1]] = (j >> 1) & 1
            c0 = c[0] | (c[1] << 1) | (c[2] << 2)
            ends[4 * axis + j] = (c0, c0 | (1 << axis))
    return ends


EDGE_ENDS = _edge_endpoints()
# the lattice point that owns edge e (offset from the cell origin) and its axis
EDGE_OWNER = np.concatenate([CORNERS[EDGE_ENDS[:, 0]], (np.arange(12) // 4)[:, None]], 1).astype(np.int32)


def _edge_between(c0, c1):
    lo, hi = min(c0, c1), max(c0, c1)
    for e in range(12):
        if EDGE_ENDS[e, 0] == lo and EDGE_ENDS[e, 1] == hi:
            return e
    raise KeyError((c0, c1))


def _faces():
    """Each face as its 4 corners in cyclic order, counter-clockwise seen from outside."""
    faces = []
    for axis in range(3):
        u, v = [(1, 2), (2, 0), (0, 1)][axis]  # u x v = +axis
        for side in (0, 1):
            ring = []
            for (a, b) in ((0, 0), (1, 0), (1, 1), (0, 1)):
                c = [0, 0, 0]
                c[axis], c[u], c[v] = side, a, b
                ring.append(c[0] | (c[1] << 1) | (c[2] << 2))
            if side == 0:  # outward normal is -axis: reverse to keep CCW from outside
                ring = ring[::-1]
            faces.append(ring)
    return faces


@functools.lru_cache(maxsize=None)
def tables():
    """Returns (edge_mask[256] uint16, tri_table[256,16] int8 padded with -1, n_tri[256] uint8)."""
    faces = _faces()
    edge_mask = np.zeros(256, np.uint16)
    tri_table = -np.ones((256, 16), np.int8)
    n_tri = np.zeros(256, np.uint8)
    for case in range(256):
        inside = [(case >> i) & 1 for i in range(8)]
        nxt = {}
        for ring in faces:
            # walk the ring; a directed segment enters at the edge where we step
            # outside->inside and leaves where we step inside->outside, which keeps the
            # inside region on the left when seen from outside the cube.
            ins, outs = [], []
            for k in range(4):
                a, b = ring[k], ring[(k + 1) % 4]
                if inside[a] != inside[b]:
                    (ins if inside[b] else outs).append((k, _edge_between(a, b)))
            if len(ins) == 1:
                nxt[outs[0][1]] = ins[0][1]
            elif len(ins) == 2:
                # ambiguous face: cut off each inside corner separately -> pair the
                # entering edge (k) with the leaving edge right after it (k+1).
                for k_in, e_in in ins:
                    e_out = [e for (k, e) in outs if k == (k_in + 1) % 4][0]
                    nxt[e_out] = e_in
        for e in nxt:
            edge_mask[case] |= np.uint16(1 << e)
        seen, tris = set(), []
        for start in sorted(nxt):
            if start in seen:
                continue
            loop, e = [], start
            while e not in seen:
                seen.add(e)
                loop.append(e)
                e = nxt[e]
            for k in range(1, len(loop) - 1):
                tris.append((loop[0], loop[k + 1], loop[k]))  # wound so that normals leave the inside
        assert len(tris) <= 5, (case, tris)
        n_tri[case] = len(tris)
        flat = [v for t in tris for v in t]
        tri_table[case, :len(flat)] = flat
    return edge_mask, tri_table, n_tri
