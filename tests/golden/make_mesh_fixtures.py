"""Finite-volume connectivity fixtures of the reference's cube meshes (BASELINE config 5).

Run HERE (the container that has /root/reference); the GPU box does not.  Reads the text siblings of the reference's
mesh files -- Gmsh 4.1 `.msh` of `meshes/3DTetrahedra_Kershaw/3DKershawTetra1` (the Kershaw family, tetrahedrised) and of
`meshes/3DHexaèdres/mesh_hexa_3`, `mesh_hexa_4` -- and stores what a cell-centred upwind finite-volume assembly needs
(what SOLVERLAB's Mesh/Cell/Face give reference src/TransportEquation.cxx:75-133): cell centres and volumes, and per
interior face the two cells and the area vector pointing from the first to the second.  The polyhedral Kershaw meshes
of BASELINE config 5 proper, `meshes/3DKershaw/Kershaw{1,2}.med`, exist as MED (HDF5) only: they are read with the
package's own minimal HDF5 reader (circulantpreconditioner_b200/hdf5_min.py, med.py -- no HDF5 / MEDfile library in this
image); the same reader applied to `mesh_hexa_3.med` / `3DKershawTetra1.med` reproduces the Gmsh-derived fixtures
(tests/test_med_reader.py).

    python tests/golden/make_mesh_fixtures.py          -> tests/golden/mesh_*.npz
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

REF = "/root/reference/meshes"
OUT = os.path.dirname(os.path.abspath(__file__))

TET_FACES = [(0, 2, 1), (0, 1, 3), (1, 2, 3), (0, 3, 2)]
HEX_FACES = [(0, 3, 2, 1), (4, 5, 6, 7), (0, 1, 5, 4), (1, 2, 6, 5), (2, 3, 7, 6), (3, 0, 4, 7)]     # Gmsh node ordering


def read_msh22(tok):
    i = tok.index("$Nodes") + 1
    nn = int(tok[i])
    rows = [tok[i + 1 + k].split() for k in range(nn)]
    index = {int(r[0]): k for k, r in enumerate(rows)}
    xyz = np.asarray([[float(v) for v in r[1:4]] for r in rows])
    i = tok.index("$Elements") + 1
    ne = int(tok[i])
    cells, kind = [], None
    for k in range(ne):
        f = [int(v) for v in tok[i + 1 + k].split()]
        if f[1] in (4, 5):
            assert kind in (None, f[1])
            kind = f[1]
            cells.append([index[v] for v in f[3 + f[2]:]])
    return xyz, np.asarray(cells), kind


def merge_duplicate_nodes(xyz, cells):
    """Nodes repeated at the same coordinates (3DKershawTetra1 has 3865 tags for 2697 points) would hide shared faces."""
    _, first, inverse = np.unique(np.round(xyz, 10), axis=0, return_index=True, return_inverse=True)
    return xyz[first], inverse.reshape(-1)[cells]


def read_msh(path):
    """Nodes and the volume elements (tets: type 4, hexahedra: type 5) of a Gmsh 2.2 / 4.1 ASCII file."""
    tok = [t.strip() for t in open(path).read().split("\n")]
    if tok[1].startswith("2."):
        return read_msh22(tok)
    i = tok.index("$Nodes") + 1
    nblocks, nnodes = [int(v) for v in tok[i].split()[:2]]
    i += 1
    tags, xyz = [], []
    for _ in range(nblocks):
        _, _, parametric, nb = [int(v) for v in tok[i].split()]
        assert parametric == 0
        i += 1
        tags += [int(tok[i + k]) for k in range(nb)]
        i += nb
        xyz += [[float(v) for v in tok[i + k].split()] for k in range(nb)]
        i += nb
    assert len(tags) == nnodes
    index = {t: k for k, t in enumerate(tags)}
    xyz = np.asarray(xyz)
    i = tok.index("$Elements") + 1
    nblocks = int(tok[i].split()[0])
    i += 1
    cells, kind = [], None
    for _ in range(nblocks):
        _, _, etype, nb = [int(v) for v in tok[i].split()]
        i += 1
        if etype in (4, 5):
            assert kind in (None, etype)
            kind = etype
            cells += [[index[int(v)] for v in tok[i + k].split()[1:]] for k in range(nb)]
        i += nb
    return xyz, np.asarray(cells), kind


def fv_geometry(xyz, cells, kind):
    faces_of = TET_FACES if kind == 4 else HEX_FACES
    nc = len(cells)
    centre = xyz[cells].mean(axis=1)                       # vertex average: inside every convex cell
    seen = {}
    fc, fa = [], []
    vol = np.zeros(nc)
    surf = np.zeros(nc)
    for c in range(nc):
        for f in faces_of:
            nodes = cells[c][list(f)]
            p = xyz[nodes]
            if len(f) == 3:
                area = 0.5 * np.cross(p[1] - p[0], p[2] - p[0])
            else:
                area = 0.5 * np.cross(p[2] - p[0], p[3] - p[1])
            fcen = p.mean(axis=0)
            if np.dot(area, fcen - centre[c]) < 0:
                area = -area                               # outward from c
            vol[c] += np.dot(fcen, area) / 3.0             # divergence theorem
            surf[c] += np.linalg.norm(area)
            key = tuple(sorted(int(v) for v in nodes))
            if key in seen:
                c0 = seen.pop(key)
                fc.append((c0, c))
                fa.append(-area)                           # from c0 towards c
            else:
                seen[key] = c
    return centre, vol, surf, np.asarray(fc, dtype=np.int32), np.asarray(fa), len(seen)


def med_fixtures():
    from circulantpreconditioner_b200 import med
    for name, rel in (("kershaw1", "3DKershaw/Kershaw1.med"), ("kershaw2", "3DKershaw/Kershaw2.med")):
        path = os.path.join(REF, rel)
        xyz, cells = med.read_med_mesh(path)
        centre, vol, surf, fc, fa, nborder = med.fv_geometry(xyz, cells)
        lo, hi = xyz.min(axis=0), xyz.max(axis=0)
        assert abs(vol.sum() - np.prod(hi - lo)) < 1e-12 * np.prod(hi - lo), (name, vol.sum())
        np.savez_compressed(os.path.join(OUT, f"mesh_{name}.npz"), centre=centre, volume=vol, surface=surf, face_cells=fc,
                            face_area=fa, bbox=np.stack([lo, hi]), source=os.path.relpath(path, "/root/reference"))
        print(f"{name}: {len(xyz)} nodes, {len(cells)} polyhedra, {len(fc)} interior faces, {nborder} border faces, "
              f"volume {vol.sum():.6f}, cell volumes {vol.min():.3e} .. {vol.max():.3e}, bbox {lo} .. {hi}")


def main():
    med_fixtures()
    jobs = [("kershaw_tetra1", os.path.join(REF, "3DTetrahedra_Kershaw", "3DKershawTetra1.msh")),
            ("hexa_3", os.path.join(REF, "3DHexaèdres", "mesh_hexa_3.msh")),
            ("hexa_4", os.path.join(REF, "3DHexaèdres", "mesh_hexa_4.msh"))]
    for name, path in jobs:
        xyz, cells, kind = read_msh(path)
        nraw = len(xyz)
        xyz, cells = merge_duplicate_nodes(xyz, cells)
        centre, vol, surf, fc, fa, nborder = fv_geometry(xyz, cells, kind)
        lo, hi = xyz.min(axis=0), xyz.max(axis=0)
        assert abs(vol.sum() - np.prod(hi - lo)) < 1e-9 * np.prod(hi - lo), (name, vol.sum())
        np.savez_compressed(os.path.join(OUT, f"mesh_{name}.npz"), centre=centre, volume=vol, surface=surf, face_cells=fc,
                            face_area=fa,
                            bbox=np.stack([lo, hi]), source=os.path.relpath(path, "/root/reference"))
        print(f"{name}: {nraw} node tags -> {len(xyz)} points, {len(cells)} cells ({'tets' if kind == 4 else 'hexahedra'}), {len(fc)} interior faces, "
              f"{nborder} border faces, volume {vol.sum():.6f}, bbox {lo} .. {hi}")


if __name__ == "__main__":
    sys.exit(main())
