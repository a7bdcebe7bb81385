"""MED mesh ingestion for BASELINE config 5: the reference's driver reads its meshes with
SOLVERLAB's `Mesh(filename)` (tests/TransportEquation_SphericalExplosion_impl_mpi.cxx:221-255), i.e. MEDCoupling on
MEDfile on HDF5 -- none of which exists in this image.  This module reads the same `.med` files through hdf5_min.py and
derives what the cell-centred upwind assembly asks of SOLVERLAB's Mesh / Cell / Face (src/TransportEquation.cxx:75-133):
cell volumes and centres, and per interior face its two cells and area vector.  Harness code around the hot path.

MED 3.x / 4.x layout (as found in the reference's files):
  /ENS_MAA/<mesh>/<step>/NOE/COO            node coordinates, NOT interlaced: x_0..x_{n-1}, y_0.., z_0..
  /ENS_MAA/<mesh>/<step>/MAI/TE4|HE8/NOD    fixed-size cells, NOT interlaced (node 0 of every cell, then node 1, ...)
  /ENS_MAA/<mesh>/<step>/MAI/POE/IFN        polyhedra: cell -> first face, 1-based, n_cells + 1 entries
  /ENS_MAA/<mesh>/<step>/MAI/POE/INN        face -> first node in NOD, 1-based, n_faces + 1 entries
  /ENS_MAA/<mesh>/<step>/MAI/POE/NOD        node numbers of the faces, 1-based
  /ENS_MAA/<mesh>/<step>/MAI/<geo>/FAM      family number of every entity (faces: TR3, QU4, POG; cells: TE4, HE8, POE)
  /FAS/<mesh>/ELEME/<family>                attribute NUM = family number; GRO/NOM = its group names (80 chars each)
"""
from __future__ import annotations

import numpy as np

from . import hdf5_min

# node sets of the faces of MED's TETRA4 / HEXA8 (orientation is fixed afterwards from the geometry)
_TET_FACES = [(0, 1, 2), (0, 3, 1), (1, 3, 2), (2, 3, 0)]
_HEX_FACES = [(0, 1, 2, 3), (4, 7, 6, 5), (0, 4, 5, 1), (1, 5, 6, 2), (2, 6, 7, 3), (3, 7, 4, 0)]


def read_med_mesh(path):
    """(xyz [n_nodes, 3], cells) of the first mesh of a MED file; cells = list of cells, a cell = list of faces, a face =
    tuple of 0-based node numbers.  3-D cells only (TE4, HE8, POE)."""
    f = hdf5_min.File(path)
    meshes = f["ENS_MAA"]
    name = meshes.keys()[0]
    steps = meshes[name]
    step = steps[steps.keys()[0]]
    coo = step["NOE/COO"].read()
    dim = int(meshes[name].attrs.get("ESP", 3))
    xyz = coo.reshape(dim, -1).T.copy()
    cells = []
    mai = step["MAI"]
    for geo, faces_of, nn in (("TE4", _TET_FACES, 4), ("HE8", _HEX_FACES, 8)):
        if geo in mai:
            nod = mai[geo]["NOD"].read().astype(np.int64).reshape(nn, -1).T - 1
            for c in nod:
                cells.append([tuple(int(c[i]) for i in fc) for fc in faces_of])
    if "POE" in mai:
        ifn = mai["POE/IFN"].read().astype(np.int64) - 1
        inn = mai["POE/INN"].read().astype(np.int64) - 1
        nod = mai["POE/NOD"].read().astype(np.int64) - 1
        for c in range(len(ifn) - 1):
            cells.append([tuple(int(v) for v in nod[inn[fc]:inn[fc + 1]]) for fc in range(ifn[c], ifn[c + 1])])
    if not cells:
        raise ValueError(f"{path}: no 3-D cells (TE4, HE8 or POE)")
    return xyz, cells


def merge_duplicate_nodes(xyz, cells):
    """Nodes repeated at the same coordinates would hide shared faces (the tetrahedrised Kershaw files have them)."""
    _, first, inverse = np.unique(np.round(xyz, 10), axis=0, return_index=True, return_inverse=True)
    inverse = inverse.reshape(-1)
    return xyz[first], [[tuple(int(inverse[v]) for v in fc) for fc in cell] for cell in cells]


def _face_fan(p):
    """Vector area, measure and the fan triangles (centroid, area vector) of a possibly non-planar polygon: triangles
    from the vertex average to every edge, as MEDCoupling triangulates polyhedron faces."""
    g = p.mean(axis=0)
    a = p - g
    b = np.roll(p, -1, axis=0) - g
    tri_area = 0.5 * np.cross(a, b)
    tri_cen = (p + np.roll(p, -1, axis=0) + g) / 3.0
    return tri_area.sum(axis=0), float(np.linalg.norm(tri_area, axis=1).sum()), tri_cen, tri_area


def fv_geometry(xyz, cells):
    """Cell centres of mass, volumes, surfaces, and per interior face (cell0, cell1) with the vector area pointing from
    cell0 to cell1; also the number of border faces.  Faces are matched by their node sets.  Volumes by the divergence
    theorem over the fan triangles (exact for the piecewise-planar cell both neighbours agree on); flux areas are the
    vector areas, so every closed cell has zero net area exactly."""
    nc = len(cells)
    centre = np.zeros((nc, 3))
    vol = np.zeros(nc)
    surf = np.zeros(nc)
    seen = {}
    fc, fa = [], []
    for c, cell in enumerate(cells):
        nodes = sorted({v for f in cell for v in f})
        g = xyz[nodes].mean(axis=0)                         # apex of the tetrahedra; inside a star-shaped cell
        m1 = np.zeros(3)
        net = np.zeros(3)
        for f in cell:
            area, measure, tcen, tarea = _face_fan(xyz[list(f)])
            sign = 1.0 if np.dot(area, xyz[list(f)].mean(axis=0) - g) >= 0 else -1.0      # outward from c
            area, tarea = sign * area, sign * tarea
            tvol = np.einsum("ij,ij->i", tcen - g, tarea) / 3.0     # tetrahedra (g, triangle)
            vol[c] += tvol.sum()
            m1 += (tvol[:, None] * (0.75 * tcen + 0.25 * g)).sum(axis=0)
            surf[c] += measure
            net += area
            key = tuple(sorted(f))
            if key in seen:
                c0 = seen.pop(key)
                fc.append((c0, c))
                fa.append(-area)
            else:
                seen[key] = c
        centre[c] = m1 / vol[c]
        if np.linalg.norm(net) > 1e-12 * surf[c]:
            raise ValueError(f"cell {c} is not closed (net area {net})")
    return centre, vol, surf, np.asarray(fc, dtype=np.int32), np.asarray(fa), len(seen)


def read_med_families(path):
    """({family number: [group names]}, {geometry: family number per entity}) of the first mesh of a MED file -- what
    SOLVERLAB's `Face::getGroupName()` is made of (the reference's assemblies branch on it: src/WaveSystem.cxx:150-171
    treats every border face outside the groups "Periodic" / "Neumann" as a wall)."""
    f = hdf5_min.File(path)
    name = f["ENS_MAA"].keys()[0]
    steps = f["ENS_MAA"][name]
    mai = steps[steps.keys()[0]]["MAI"]
    families = {0: []}
    fas = f["FAS"][name] if name in f["FAS"] else None
    if fas is not None and "ELEME" in fas:
        for fam in fas["ELEME"].keys():
            g = fas["ELEME"][fam]
            names = []
            if "GRO" in g:
                raw = g["GRO/NOM"].read().astype(np.uint8)
                names = [bytes(row).split(b"\0")[0].decode("latin-1").strip() for row in raw.reshape(-1, raw.shape[-1])]
            families[int(g.attrs["NUM"])] = names
    per_geo = {geo: mai[geo]["FAM"].read().astype(np.int64) for geo in mai.keys() if "FAM" in mai[geo]}
    return families, per_geo
