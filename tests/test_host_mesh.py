"""Marching-cubes case table (anerf_b200/mc_table.py): structural checks on all 256 cases, the committed C header, and
the PLY writer.  No GPU."""
import os

import numpy as np

from anerf_b200 import mc_table as mt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_committed_header_is_the_generated_table():
    with open(os.path.join(ROOT, "anerf_b200", "csrc", "mc_table.inc")) as fh:
        assert fh.read() == mt.emit_header()


def test_every_case_is_a_set_of_closed_oriented_loops():
    assert mt.MAX_TRIS == 5 and mt.TRI_COUNT[0] == 0 and mt.TRI_COUNT[255] == 0
    for case in range(256):
        inside = [(case >> c) & 1 for c in range(8)]
        crossed = {e for e, (a, b) in enumerate(mt.EDGE_CORNERS) if inside[a] != inside[b]}
        tris = mt.TRI_TABLE[case, :mt.TRI_COUNT[case]]
        assert set(tris.reshape(-1).tolist()) == crossed                     # exactly the crossed edges carry vertices
        assert (mt.TRI_TABLE[case, mt.TRI_COUNT[case]:] == -1).all()
        # directed edges: interior diagonals appear once in each direction, boundary (cube-face) segments once
        d = {}
        for a, b, c in tris.tolist():
            assert len({a, b, c}) == 3
            for u, v in ((a, b), (b, c), (c, a)):
                d[(u, v)] = d.get((u, v), 0) + 1
        assert all(n == 1 for n in d.values())
        boundary = [(u, v) for (u, v) in d if (v, u) not in d]
        # every crossed cube edge is the end of exactly two boundary segments (it lies on two cube faces)
        deg = {}
        for u, v in boundary:
            deg[u] = deg.get(u, 0) + 1
            deg[v] = deg.get(v, 0) + 1
        assert all(deg.get(e, 0) == 2 for e in crossed), case
    # complementary cases use the same edges
    for case in range(256):
        assert set(mt.TRI_TABLE[case][mt.TRI_TABLE[case] >= 0].tolist()) == set(mt.TRI_TABLE[255 - case][mt.TRI_TABLE[255 - case] >= 0].tolist())


def test_ply_writer_round_trip(tmp_path):
    import torch
    from anerf_b200 import mesh
    v = torch.tensor([[0., 0., 0.], [1., 0., 0.], [0., 1., 0.], [0., 0., 1.]])
    f = torch.tensor([[0, 2, 1], [0, 1, 3], [1, 2, 3], [0, 3, 2]])
    p = tmp_path / "t.ply"
    mesh.export_ply(str(p), v, f)
    raw = p.read_bytes()
    head, body = raw.split(b"end_header\n")
    assert b"element vertex 4" in head and b"element face 4" in head and b"binary_little_endian" in head
    vv = np.frombuffer(body[:48], dtype="<f4").reshape(4, 3)
    rec = np.frombuffer(body[48:], dtype=[("n", "u1"), ("idx", "<i4", (3,))])
    assert np.array_equal(vv, v.numpy()) and (rec["n"] == 3).all() and np.array_equal(rec["idx"], f.numpy())
