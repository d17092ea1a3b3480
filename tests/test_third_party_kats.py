"""CPU known-answer tests for the three third-party algorithms the oracle RESTATES (their sources are not under
/root/reference, so these rows stay "parity unpinned" against the real packages): torchsparse v1.4.0 sparse convolution
rules, inplace_abn's |gamma| + eps scaling, PyMCubes' marching cubes.  Every expected value below was worked out by hand
from the rule stated in the test, not produced by the code under test; the GPU tests then compare the CUDA kernels with
these same oracle functions (tests/test_gpu_parity.py), so a silent change of a rule on either side shows up.

Upstream rules restated (torchsparse tag v1.4.0: torchsparse/nn/utils/kernel.py `get_kernel_offsets`,
torchsparse/nn/functional/downsample.py `spdownsample`, torchsparse/nn/functional/conv.py `conv3d`; inplace_abn
inplace_abn/abn.py `InPlaceABN.forward` -> `weight.abs() + eps`; PyMCubes mcubes/src/marchingcubes.h)."""
import numpy as np
import torch

from oracle import recon_oracle as O


def test_kernel_offsets_x_fastest():
    off = O.kernel_offsets(1)
    assert off.shape == (27, 3)
    assert off[0].tolist() == [-1, -1, -1] and off[1].tolist() == [0, -1, -1] and off[2].tolist() == [1, -1, -1]   # x runs fastest
    assert off[3].tolist() == [-1, 0, -1] and off[9].tolist() == [-1, -1, 0] and off[13].tolist() == [0, 0, 0]
    assert O.kernel_offsets(2)[26].tolist() == [2, 2, 2]                                                               # scaled by the tensor stride


def test_stride2_output_coordinates_rule():
    """spdownsample, kernel (3) != stride (2): candidates = input + every offset, kept when every component is a multiple of
    2 * tensor_stride and >= the per-axis minimum of the inputs; unique, sorted."""
    # one voxel at (3,3,3): candidates are {2,3,4}^3, the even ones {2,4}^3, all < min 3 excluded -> only (4,4,4)
    out = O.downsample_coords(torch.tensor([[3, 3, 3]]), 1)
    assert out.tolist() == [[4, 4, 4]]
    # one voxel at (2,2,2): candidates {1,2,3}^3 -> even and >= 2 -> (2,2,2)
    assert O.downsample_coords(torch.tensor([[2, 2, 2]]), 1).tolist() == [[2, 2, 2]]
    # two voxels (0,0,0) and (1,0,0): candidates x in {-1..2}; even and >= 0: x in {0, 2}, y = z = 0
    assert O.downsample_coords(torch.tensor([[0, 0, 0], [1, 0, 0]]), 1).tolist() == [[0, 0, 0], [2, 0, 0]]
    # tensor stride 2 (second down-conv): multiples of 4
    assert O.downsample_coords(torch.tensor([[2, 2, 2]]), 2).tolist() == [[4, 4, 4]]
    assert O.downsample_coords(torch.tensor([[0, 4, 8]]), 2).tolist() == [[0, 4, 8]]


def test_sparse_conv_stride1_and_stride2_by_hand():
    """out[j] = sum_k W[k]^T in[i] over the inputs i at coord_j + offset_k (no spatial flip, kernel index = offset index)."""
    coords = torch.tensor([[0, 0, 0, 0], [1, 0, 0, 0]], dtype=torch.int32)          # (x, y, z, batch)
    feats = torch.tensor([[1.0], [10.0]])
    W = torch.arange(27, dtype=torch.float32).reshape(27, 1, 1) + 1.0                 # W[k] = k + 1
    out, oc, stride, cm, km = O.torchsparse_conv3d(feats, coords, 1, W, 1, False, {}, {})
    # output 0 at (0,0,0): itself via offset (0,0,0) = k 13 (weight 14) and the neighbour (1,0,0) via offset (+1,0,0) = k 14 (15)
    # output 1 at (1,0,0): itself via k 13 and the neighbour (0,0,0) via offset (-1,0,0) = k 12 (weight 13)
    assert torch.equal(oc, coords) and stride == 1
    assert out[:, 0].tolist() == [1 * 14 + 10 * 15, 10 * 14 + 1 * 13]
    # stride 2: outputs at (0,0,0) and (2,0,0) (previous test); (2,0,0) sees only (1,0,0) through offset (-1,0,0) = k 12
    out2, oc2, stride2, cm2, km2 = O.torchsparse_conv3d(feats, coords, 1, W, 2, False, {}, {})
    assert oc2[:, :3].tolist() == [[0, 0, 0], [2, 0, 0]] and stride2 == 2
    assert out2[:, 0].tolist() == [1 * 14 + 10 * 15, 10 * 13]
    # transposed conv: the cached map of the down-conv with in / out swapped, same kernel index, fine coordinates restored
    up, uc, ustride, _, _ = O.torchsparse_conv3d(torch.tensor([[2.0], [3.0]]), oc2, 2, W, 2, True, cm2, km2)
    assert torch.equal(uc, coords) and ustride == 1
    # fine (0,0,0) <- coarse (0,0,0) via k 13; fine (1,0,0) <- coarse (0,0,0) via k 14 and coarse (2,0,0) via k 12
    assert up[:, 0].tolist() == [2 * 14, 2 * 15 + 3 * 13]


def test_inplace_abn_uses_abs_gamma_plus_eps():
    x = torch.tensor([[[[1.0, 3.0]]], [[[5.0, 7.0]]]])                                # [N=2, C=1, 1, 2]: mean 4, biased var 5
    y_pos = O.inplace_abn(x, torch.tensor([2.0]), torch.tensor([0.5]))
    y_neg = O.inplace_abn(x, torch.tensor([-2.0]), torch.tensor([0.5]))
    assert torch.equal(y_pos, y_neg)                                                  # the sign of gamma is dropped
    z = (x - 4.0) / np.sqrt(5.0 + 1e-5) * (2.0 + 1e-5) + 0.5
    want = torch.where(z >= 0, z, 0.01 * z)                                           # leaky ReLU 0.01
    assert float((y_pos - want).abs().max()) < 1e-6
    y_bn = torch.nn.functional.leaky_relu(torch.nn.functional.batch_norm(x, None, None, torch.tensor([-2.0]), torch.tensor([0.5]),
                                                                         True, 0.1, 1e-5), 0.01)
    assert float((y_neg - y_bn).abs().max()) > 1.0                                    # plain BatchNorm would flip the sign


def test_marching_cubes_single_cell_cases():
    """One cell (2 x 2 x 2 samples): the vertex set is exactly the sign-changing edges, positions by linear interpolation,
    and the triangle count follows the case class of the classic 256-entry table."""
    def cell(inside):
        u = -np.ones((2, 2, 2))
        for c in inside:
            u[c] = 1.0
        return u
    # corner (0,0,0) inside: three crossing edges at t = 0.5 on each axis, one triangle
    v, t, case = O.marching_cubes(cell([(0, 0, 0)]), 0.0)
    assert len(v) == 3 and len(t) == 1 and int(case[0, 0, 0]) == 1
    assert sorted(map(tuple, v.tolist())) == [(0.0, 0.0, 0.5), (0.0, 0.5, 0.0), (0.5, 0.0, 0.0)]
    # interpolation: u = +3 inside, -1 outside -> crossing at 3 / 4 of the edge
    u = cell([(0, 0, 0)])
    u[0, 0, 0] = 3.0
    v, _, _ = O.marching_cubes(u, 0.0)
    assert sorted(map(tuple, v.tolist())) == [(0.0, 0.0, 0.75), (0.0, 0.75, 0.0), (0.75, 0.0, 0.0)]
    # one whole face inside (x = 0 plane): 4 crossing edges, a quad = 2 triangles
    v, t, _ = O.marching_cubes(cell([(0, 0, 0), (0, 0, 1), (0, 1, 0), (0, 1, 1)]), 0.0)
    assert len(v) == 4 and len(t) == 2 and np.allclose(v[:, 0], 0.5)
    # two diagonally opposite corners: 6 crossing edges, two separate triangles
    v, t, _ = O.marching_cubes(cell([(0, 0, 0), (1, 1, 1)]), 0.0)
    assert len(v) == 6 and len(t) == 2
    # complement: the same edges cross when inside / outside are swapped (the triangulation may differ: a face with two
    # diagonal inside corners is ambiguous, and the table separates the inside corners -- 2 triangles -- but joins the
    # outside ones -- 4 triangles)
    for inside in ([(0, 0, 0)], [(0, 0, 0), (1, 0, 0)], [(0, 0, 0), (1, 1, 0)], [(0, 0, 0), (0, 1, 1), (1, 0, 1)]):
        a, ta, _ = O.marching_cubes(cell(inside), 0.0)
        b, tb, _ = O.marching_cubes(-cell(inside), 0.0)
        assert sorted(map(tuple, np.round(a, 6).tolist())) == sorted(map(tuple, np.round(b, 6).tolist()))
    assert len(O.marching_cubes(cell([(0, 0, 0), (1, 1, 0)]), 0.0)[1]) == 2 and len(O.marching_cubes(-cell([(0, 0, 0), (1, 1, 0)]), 0.0)[1]) == 4
    # every one of the 256 cases: triangles only use crossing edges, and every crossing edge is used
    from o2345 import mc_tables as T
    _, tri_table, n_tri = T.tables()
    for c in range(256):
        ins = [(c >> i) & 1 for i in range(8)]
        crossing = {e for e, (a, b) in enumerate(T.EDGE_ENDS) if ins[a] != ins[b]}
        used = {int(e) for e in tri_table[c, :3 * int(n_tri[c])]}
        assert used == crossing, c
