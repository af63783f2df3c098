"""Case table of marching cubes, GENERATED (not transcribed): for each of the 256 sign configurations of a cube's
corners, the triangles of the isosurface patch as triples of cube-edge indices.

Construction (the classic polygonisation rules): on every cube face the crossed edges are joined pairwise -- a face with
two crossed edges gets one segment, an ambiguous face (four crossed edges, inside corners on a diagonal) gets two
segments that cut the INSIDE corners off.  That choice depends only on the face's own four corners, so the two cubes
sharing a face always agree and the surface is watertight for any input.  Every crossed edge lies on two faces, so the
segments form closed loops; each loop is fan-triangulated and oriented so that its normal points from inside (value >
iso) to outside.

The reference extracts meshes with the third-party PyMCubes (`mcubes.marching_cubes`, run_render.py:983-986; not
installed here, version unpinned): same vertices (linear interpolation on the crossed cube edges), same surface up to
the resolution of ambiguous faces and the choice of diagonals inside a patch -- parity with that package is UNPINNED
(SURVEY.md 8(c)); tests check the geometry instead (closed 2-manifold, Euler characteristic, area/volume, orientation).

Corner c has offset (c & 1, (c >> 1) & 1, (c >> 2) & 1) along the volume's (i, j, k) axes; edge e joins EDGE_CORNERS[e].
"""
import numpy as np

CORNER_OFFSETS = np.array([[c & 1, (c >> 1) & 1, (c >> 2) & 1] for c in range(8)], dtype=np.int64)
# 12 edges: 4 along i, 4 along j, 4 along k
EDGE_CORNERS = np.array([(a, a | (1 << ax)) for ax in range(3) for a in range(8) if not a & (1 << ax)], dtype=np.int64)
EDGE_AXIS = np.repeat(np.arange(3), 4)
_EDGE_OF = {(int(a), int(b)): e for e, (a, b) in enumerate(EDGE_CORNERS)}


def _face_cycles():
    """The 6 faces as cycles of 4 corners (neighbours along the cycle share a cube edge)."""
    faces = []
    for ax in range(3):
        u, v = [1 << b for b in range(3) if b != ax]
        for side in (0, 1 << ax):
            faces.append([side, side | u, side | u | v, side | v])
    return faces


def _edge(a, b):
    return _EDGE_OF[(min(a, b), max(a, b))]


def _field(inside, p):
    """Trilinear interpolation of the +-1 corner indicator at p in [0,1]^3 (orientation test only)."""
    f = 0.0
    for c in range(8):
        w = 1.0
        for ax in range(3):
            w *= p[ax] if (c >> ax) & 1 else 1.0 - p[ax]
        f += w * (1.0 if inside[c] else -1.0)
    return f


def build_tables():
    faces = _face_cycles()
    mid = (CORNER_OFFSETS[EDGE_CORNERS[:, 0]] + CORNER_OFFSETS[EDGE_CORNERS[:, 1]]) * 0.5
    tri_lists = []
    for case in range(256):
        inside = [(case >> c) & 1 == 1 for c in range(8)]
        nbr = {}
        for cyc in faces:
            crossed = [(cyc[i], cyc[(i + 1) % 4]) for i in range(4) if inside[cyc[i]] != inside[cyc[(i + 1) % 4]]]
            if len(crossed) == 2:
                pairs = [(_edge(*crossed[0]), _edge(*crossed[1]))]
            elif len(crossed) == 4:
                # cut each inside corner off: its two face edges are joined
                pairs = []
                for i in range(4):
                    if inside[cyc[i]]:
                        pairs.append((_edge(cyc[i], cyc[(i + 1) % 4]), _edge(cyc[i], cyc[(i - 1) % 4])))
            else:
                pairs = []
            for a, b in pairs:
                nbr.setdefault(a, []).append(b)
                nbr.setdefault(b, []).append(a)
        assert all(len(v) == 2 for v in nbr.values()), case
        tris, seen = [], set()
        for start in sorted(nbr):
            if start in seen:
                continue
            loop, prev, cur = [start], None, start
            seen.add(start)
            while True:
                nxt = [x for x in nbr[cur] if x != prev]
                nxt = nxt[0] if nxt else nbr[cur][0]
                if nbr[cur][0] == nbr[cur][1]:
                    nxt = nbr[cur][0]
                if nxt == start:
                    break
                loop.append(nxt)
                seen.add(nxt)
                prev, cur = cur, nxt
            assert len(loop) >= 3, (case, loop)
            # fan apex: avoid diagonals whose two ends lie on one cube face (the neighbouring cube may create the same
            # edge there, which would make it non-manifold); all loops of a cube admit such an apex or minimise the count
            def on_common_face(e1, e2):
                c1, c2 = set(map(int, EDGE_CORNERS[e1])), set(map(int, EDGE_CORNERS[e2]))
                return any(c1 <= set(f) and c2 <= set(f) for f in faces)
            best = None
            for r in range(len(loop)):
                lp = loop[r:] + loop[:r]
                bad = sum(on_common_face(lp[0], lp[i]) for i in range(2, len(lp) - 1))
                if best is None or bad < best[0]:
                    best = (bad, lp)
            loop = best[1]
            fan = [(loop[0], loop[i], loop[i + 1]) for i in range(1, len(loop) - 1)]
            # orientation: normal from inside to outside (the indicator decreases along it)
            n = np.zeros(3)
            for a, b, c in fan:
                n += np.cross(mid[b] - mid[a], mid[c] - mid[a])
            g = mid[loop].mean(0)
            n = n / max(np.linalg.norm(n), 1e-12)
            if _field(inside, g + 0.05 * n) > _field(inside, g - 0.05 * n):
                fan = [(a, c, b) for a, b, c in fan]
            tris += fan
        tri_lists.append(tris)
    max_tris = max(len(t) for t in tri_lists)
    table = -np.ones((256, max_tris, 3), dtype=np.int32)
    counts = np.zeros(256, dtype=np.int32)
    for case, tris in enumerate(tri_lists):
        counts[case] = len(tris)
        for i, t in enumerate(tris):
            table[case, i] = t
    return counts, table


TRI_COUNT, TRI_TABLE = build_tables()
MAX_TRIS = TRI_TABLE.shape[1]


def emit_header():
    """C tables for csrc/mesh_kernels.cuh (committed as csrc/mc_table.inc; tests/test_host_mesh.py checks that the
    committed file equals this output)."""
    lines = ["// GENERATED by `python -m anerf_b200.mc_table` (anerf_b200/mc_table.py) -- do not edit.",
             f"#define ANERF_MC_MAX_TRIS {MAX_TRIS}",
             "static const signed char kMcTriCount[256] = {" + ", ".join(str(int(x)) for x in TRI_COUNT) + "};",
             f"static const signed char kMcTriTable[256][{MAX_TRIS * 3}] = {{"]
    for case in range(256):
        lines.append("  {" + ", ".join(str(int(x)) for x in TRI_TABLE[case].reshape(-1)) + "},")
    lines.append("};")
    lines.append("// edge e joins corners kMcEdgeCorner[e][0] < kMcEdgeCorner[e][1]; corner c sits at (c&1, (c>>1)&1, (c>>2)&1)")
    lines.append("static const signed char kMcEdgeCorner[12][2] = {" + ", ".join("{%d, %d}" % (int(a), int(b)) for a, b in EDGE_CORNERS) + "};")
    lines.append("static const signed char kMcEdgeAxis[12] = {" + ", ".join(str(int(x)) for x in EDGE_AXIS) + "};")
    return "\n".join(lines) + "\n"


if __name__ == "__main__":
    import os
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "mc_table.inc")
    with open(out, "w") as fh:
        fh.write(emit_header())
    print("wrote", out)
