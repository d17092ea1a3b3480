"""Mesh tail of the reconstruction path without trimesh (not installed here).

Restates, on plain numpy arrays:
  * `merge_vertices`        what `trimesh.Trimesh(vertices, faces, vertex_colors=...)` does with its default `process=True`
                            before reference reconstruction/models/trainer_generic.py:1374-1380 exports mesh.ply: vertices
                            that coincide after rounding to 8 decimals are merged (first occurrence kept, with its colour),
                            faces are re-indexed, vertices no face references are dropped.  trimesh's source is not under
                            /root/reference: PARITY UNPINNED (only the vertex ORDER could differ; geometry cannot).
  * `write_ply`             binary little-endian PLY with per-vertex RGBA, the layout trimesh emits.
  * `to_viewer_frame` + `write_obj` / `write_glb`
                            reference utils/utils.py:31-45 `convert_mesh_format`: rotate +90 degrees about x, 180 degrees
                            about z, negate x, reverse the face winding -- altogether (x, y, z) -> (x, z, y) with flipped
                            faces -- then `.obj` with `v x y z r g b` lines (trimesh's include_color=True) or a binary glTF.
"""
from __future__ import annotations

import json
import struct

import numpy as np


def merge_vertices(vertices, triangles, colors=None, digits=8, candidates=None):
    """candidates (optional): indices of the only vertices that can coincide with another one (marching cubes: those on a
    lattice point).  With fewer than two of them, or no coinciding pair among them, the arrays are returned untouched; the
    sort over all vertices is only paid when something really merges."""
    v = np.asarray(vertices)
    f = np.asarray(triangles)
    if len(v) == 0 or len(f) == 0:
        return v, f, colors
    if candidates is not None:
        cand = np.asarray(candidates, np.int64)
        if cand.size < 2:
            return v, f, colors
        ck = np.round(np.asarray(v[cand], np.float64) * 10.0 ** digits).astype(np.int64)
        if len(np.unique(ck, axis=0)) == len(ck):
            return v, f, colors
    v = np.asarray(v, np.float64)
    f = np.asarray(f, np.int64).reshape(-1, 3)
    key = np.round(v * 10.0 ** digits).astype(np.int64)
    _, first, inverse = np.unique(key, axis=0, return_index=True, return_inverse=True)
    inverse = np.asarray(inverse).reshape(-1)
    # keep the FIRST occurrence of every position, in order of first occurrence
    order = np.argsort(first, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    keep = first[order]
    remap = rank[inverse]
    f2 = remap[f]
    used = np.zeros(len(keep), bool)
    used[f2.reshape(-1)] = True
    if not used.all():
        compact = np.cumsum(used) - 1
        f2 = compact[f2]
        keep = keep[used]
    return v[keep], f2, (None if colors is None else np.asarray(colors)[keep])


def write_ply(path, vertices, triangles, colors):
    """Binary little-endian PLY with per-vertex RGBA."""
    v = np.asarray(vertices, np.float32)
    f = np.asarray(triangles, np.int32)
    c = np.asarray(colors, np.uint8)
    if c.shape[1] == 3:
        c = np.concatenate([c, np.full((len(c), 1), 255, np.uint8)], 1)
    header = ("ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\n"
              "property float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nproperty uchar alpha\n"
              "element face %d\nproperty list uchar int vertex_indices\nend_header\n") % (len(v), len(f))
    vrec = np.empty(len(v), dtype=[("p", "<f4", 3), ("c", "u1", 4)])
    vrec["p"], vrec["c"] = v, c
    frec = np.empty(len(f), dtype=[("n", "u1"), ("i", "<i4", 3)])
    frec["n"], frec["i"] = 3, f
    with open(path, "wb") as fh:
        fh.write(header.encode("ascii"))
        fh.write(vrec.tobytes())
        fh.write(frec.tobytes())


def read_ply(path):
    """Reads back a file written by write_ply -> (vertices float32 [n,3], triangles int32 [m,3], colors uint8 [n,4])."""
    with open(path, "rb") as fh:
        raw = fh.read()
    end = raw.index(b"end_header\n") + len(b"end_header\n")
    head = raw[:end].decode("ascii")
    nv = int(head.split("element vertex ")[1].split("\n")[0])
    nf = int(head.split("element face ")[1].split("\n")[0])
    vrec = np.frombuffer(raw, dtype=[("p", "<f4", 3), ("c", "u1", 4)], count=nv, offset=end)
    frec = np.frombuffer(raw, dtype=[("n", "u1"), ("i", "<i4", 3)], count=nf, offset=end + nv * 16)
    return vrec["p"].copy(), frec["i"].copy(), vrec["c"].copy()


def to_viewer_frame(vertices, triangles):
    """convert_mesh_format's transform chain (reference utils/utils.py:35-41), applied exactly in its order."""
    v = np.asarray(vertices, np.float64)
    rx = np.array([[1, 0, 0], [0, 0, -1], [0, 1, 0]], np.float64)        # rotation_matrix(pi / 2, [1, 0, 0])
    rz = np.array([[-1, 0, 0], [0, -1, 0], [0, 0, 1]], np.float64)       # rotation_matrix(pi, [0, 0, 1])
    v = v @ rx.T
    v = v @ rz.T
    v[:, 0] = -v[:, 0]
    return v, np.fliplr(np.asarray(triangles)).copy()


def write_obj(path, vertices, triangles, colors):
    v = np.asarray(vertices, np.float64)
    c = np.asarray(colors, np.float64)[:, :3] / 255.0
    f = np.asarray(triangles, np.int64) + 1
    with open(path, "w") as fh:
        fh.write("# o2345-b200\n")
        for p, q in zip(v, c):
            fh.write("v %.8f %.8f %.8f %.8f %.8f %.8f\n" % (p[0], p[1], p[2], q[0], q[1], q[2]))
        for t in f:
            fh.write("f %d %d %d\n" % (t[0], t[1], t[2]))


def write_glb(path, vertices, triangles, colors):
    """Binary glTF 2.0: one mesh, POSITION float32, COLOR_0 normalised uint8 RGBA, uint32 indices."""
    v = np.ascontiguousarray(vertices, np.float32)
    c = np.asarray(colors, np.uint8)
    if c.shape[1] == 3:
        c = np.concatenate([c, np.full((len(c), 1), 255, np.uint8)], 1)
    c = np.ascontiguousarray(c)
    idx = np.ascontiguousarray(np.asarray(triangles, np.uint32).reshape(-1))
    blobs = [v.tobytes(), c.tobytes(), idx.tobytes()]
    offs, total = [], 0
    for b in blobs:
        offs.append(total)
        total += (len(b) + 3) // 4 * 4
    bin_chunk = bytearray(total)
    for o, b in zip(offs, blobs):
        bin_chunk[o:o + len(b)] = b
    doc = {
        "asset": {"version": "2.0", "generator": "o2345-b200"},
        "scene": 0, "scenes": [{"nodes": [0]}], "nodes": [{"mesh": 0}],
        "meshes": [{"primitives": [{"attributes": {"POSITION": 0, "COLOR_0": 1}, "indices": 2, "mode": 4}]}],
        "buffers": [{"byteLength": total}],
        "bufferViews": [{"buffer": 0, "byteOffset": offs[0], "byteLength": len(blobs[0]), "target": 34962},
                        {"buffer": 0, "byteOffset": offs[1], "byteLength": len(blobs[1]), "target": 34962},
                        {"buffer": 0, "byteOffset": offs[2], "byteLength": len(blobs[2]), "target": 34963}],
        "accessors": [{"bufferView": 0, "componentType": 5126, "count": int(len(v)), "type": "VEC3",
                       "min": v.min(0).tolist() if len(v) else [0, 0, 0], "max": v.max(0).tolist() if len(v) else [0, 0, 0]},
                      {"bufferView": 1, "componentType": 5121, "normalized": True, "count": int(len(c)), "type": "VEC4"},
                      {"bufferView": 2, "componentType": 5125, "count": int(len(idx)), "type": "SCALAR"}],
    }
    js = json.dumps(doc, separators=(",", ":")).encode("utf-8")
    js += b" " * ((4 - len(js) % 4) % 4)
    with open(path, "wb") as fh:
        fh.write(struct.pack("<III", 0x46546C67, 2, 12 + 8 + len(js) + 8 + len(bin_chunk)))
        fh.write(struct.pack("<II", len(js), 0x4E4F534A))
        fh.write(js)
        fh.write(struct.pack("<II", len(bin_chunk), 0x004E4942))
        fh.write(bytes(bin_chunk))


def convert_mesh_format(exp_dir, output_format=".obj"):
    """reference utils/utils.py:31-45: <exp_dir>/mesh.ply -> <exp_dir>/mesh.obj | mesh.glb in the viewer frame."""
    import os
    v, f, c = read_ply(os.path.join(exp_dir, "mesh.ply"))
    v2, f2 = to_viewer_frame(v, f)
    out = os.path.join(exp_dir, f"mesh{output_format}")
    if output_format == ".obj":
        write_obj(out, v2, f2, c)
    else:
        write_glb(out, v2, f2, c)
    return out
