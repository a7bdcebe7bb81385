"""The package's minimal HDF5 / MED reader (hdf5_min.py, med.py) against the reference's own mesh files.

BASELINE config 5 names the polyhedral Kershaw meshes, which the reference ships as MED (HDF5) only
(meshes/3DKershaw/Kershaw{1,2}.med) and reads through SOLVERLAB / MEDCoupling / MEDfile / HDF5
(tests/TransportEquation_SphericalExplosion_impl_mpi.cxx:221-255) -- none of them in this image.  The reader is pinned
two ways: (1) against the reference's mesh table (meshes/README.md: node and cell counts of every .med file it ships); (2) against the
independent Gmsh text siblings of the same meshes (tests/golden/make_mesh_fixtures.py): the MED route and the .msh route
must give the same finite-volume geometry.  Needs /root/reference (this container; the GPU box only uses the fixtures).
"""
import os

import numpy as np
import pytest

from circulantpreconditioner_b200 import hdf5_min as H
from circulantpreconditioner_b200 import med
from circulantpreconditioner_b200 import meshes as MS

REF = "/root/reference/meshes"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference is not available here")

# meshes/README.md: (file, nodes, cells)
TABLE = [("3DHexaèdres/mesh_hexa_1.med", 27, 8), ("3DHexaèdres/mesh_hexa_2.med", 125, 64),
         ("3DHexaèdres/mesh_hexa_3.med", 729, 512), ("3DHexaèdres/mesh_hexa_4.med", 4913, 4096),
         ("3DHexaèdres/mesh_hexa_5.med", 35937, 32768),
         ("3DTetrahedra/mesh_tetra_0.med", 80, 215), ("3DTetrahedra/mesh_tetra_1.med", 488, 2003),
         ("3DTetrahedra/mesh_tetra_2.med", 857, 3898), ("3DTetrahedra/mesh_tetra_3.med", 1601, 7711),
         ("3DTetrahedra/mesh_tetra_4.med", 2997, 15266), ("3DTetrahedra/mesh_tetra_5.med", 5692, 30480),
         ("3DTetrahedra/mesh_tetra_6.med", 10994, 61052),
         ("3DTetrahedra_Kershaw/3DKershawTetra1.med", 3865, 11072), ("3DTetrahedra_Kershaw/3DKershawTetra2.med", 31793, 93440),
         ("3DKershaw/Kershaw1.med", 729, 512), ("3DKershaw/Kershaw2.med", 4913, 4096)]


@needs_ref
def test_hdf5_tree_of_a_med_file():
    f = H.File(os.path.join(REF, "3DKershaw", "Kershaw1.med"))
    assert f.O == 8 and f.L == 8
    assert sorted(f.root.keys()) == ["ENS_MAA", "FAS", "INFOS_GENERALES"]
    info = f["INFOS_GENERALES"].attrs
    assert (info["MAJ"], info["MIN"]) >= (3, 0)                      # MED 3.x / 4.x layout
    mesh = f["ENS_MAA/mesh"]
    assert mesh.attrs["ESP"] == 3 and mesh.attrs["DIM"] == 3
    step = mesh[mesh.keys()[0]]
    poe = step["MAI/POE"]
    assert poe.attrs["GEO"] == 500                                   # MED_POLYHEDRON
    for name, n in (("IFN", 513), ("INN", 3073), ("NOD", 12288)):
        d = poe[name]
        assert d.shape == (n,) and d.dtype == np.dtype("<i4") and d.attrs["NBR"] == n
    ifn, inn, nod = (poe[k].read() for k in ("IFN", "INN", "NOD"))
    assert ifn[0] == 1 and ifn[-1] == 3073 and np.all(np.diff(ifn) == 6)       # six faces per Kershaw cell
    assert inn[0] == 1 and inn[-1] == 12289 and np.all(np.diff(inn) == 4)      # four nodes per face
    assert nod.min() == 1 and nod.max() == 729
    coo = step["NOE/COO"]
    assert coo.shape == (3 * 729,) and coo.dtype == np.dtype("<f8") and coo.attrs["NBR"] == 729
    with pytest.raises(KeyError):
        f["ENS_MAA/nothing"]
    # walk() reaches every object exactly once per path and lists the family groups too
    paths = [p for p, _ in f.walk()]
    assert len(paths) == len(set(paths)) and any(p.startswith("/FAS/mesh/ELEME/") for p in paths)


def test_old_style_hdf5_file_from_scipys_test_data():
    """The other flavour of HDF5 -- superblock version 0 behind a 512-byte user block, a symbol-table group, version-1
    object headers and attributes, layout message version 2 -- on the one such file this image holds: scipy's MATLAB 7.3
    sample (testdouble = linspace(0, 2 pi, 9)).  The reference's .med files exercise the new-style structures."""
    import scipy.io
    p = os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
    if not os.path.exists(p):
        pytest.skip("scipy's test data is not installed")
    f = H.File(p)
    assert f.root.keys() == ["testdouble"]
    d = f["testdouble"]
    assert d.shape == (9, 1) and d.dtype == np.dtype("<f8") and d.attrs["MATLAB_class"] == "double"
    assert np.allclose(d.read().ravel(), np.linspace(0.0, 2.0 * np.pi, 9), rtol=0, atol=1e-15)


def test_not_hdf5(tmp_path):
    p = tmp_path / "x.med"
    p.write_bytes(b"$MeshFormat\n2.2 0 8\n")
    with pytest.raises(H.Hdf5Error):
        H.File(str(p))


@needs_ref
@pytest.mark.parametrize("rel,nnodes,ncells", TABLE)
def test_counts_of_the_reference_mesh_table(rel, nnodes, ncells):
    xyz, cells = med.read_med_mesh(os.path.join(REF, rel))
    assert xyz.shape == (nnodes, 3) and len(cells) == ncells
    assert np.allclose(xyz.min(axis=0), 0.0) and np.allclose(xyz.max(axis=0), 1.0)      # the unit cube


@needs_ref
@pytest.mark.parametrize("rel,fixture", [("3DHexaèdres/mesh_hexa_3.med", "hexa_3"),
                                         ("3DTetrahedra_Kershaw/3DKershawTetra1.med", "kershaw_tetra1")])
def test_med_route_equals_gmsh_route(rel, fixture):
    """Same mesh, two file formats, two readers: identical cells, volumes, surfaces, interior faces and area vectors
    (planar faces: the fan triangulation of med.fv_geometry and the diagonal formula of the Gmsh route agree; centres
    are centres of mass here and vertex averages there, which coincide for tetrahedra and parallelepipeds)."""
    xyz, cells = med.read_med_mesh(os.path.join(REF, rel))
    xyz, cells = med.merge_duplicate_nodes(xyz, cells)
    centre, vol, surf, fc, fa, nborder = med.fv_geometry(xyz, cells)
    fix = MS.load_fixture(fixture)
    assert np.allclose(centre, fix["centre"], rtol=0, atol=1e-14)
    assert np.allclose(vol, fix["volume"], rtol=1e-12, atol=0)
    assert np.allclose(surf, fix["surface"], rtol=1e-12, atol=0)
    # interior faces as an unordered set of (cell pair -> area vector from the lower to the higher cell)
    def as_map(fc, fa):
        out = {}
        for (a, b), s in zip(fc, fa):
            out[(min(a, b), max(a, b))] = s if a < b else -s
        return out
    got, want = as_map(fc, fa), as_map(fix["face_cells"], fix["face_area"])
    assert got.keys() == want.keys()
    assert max(np.abs(got[k] - want[k]).max() for k in got) < 1e-14


@needs_ref
@pytest.mark.parametrize("name,rel,ncells,nfaces,nborder", [("kershaw1", "3DKershaw/Kershaw1.med", 512, 1344, 384),
                                                            ("kershaw2", "3DKershaw/Kershaw2.med", 4096, 11520, 1536)])
def test_kershaw_fixtures_regenerate(name, rel, ncells, nfaces, nborder):
    xyz, cells = med.read_med_mesh(os.path.join(REF, rel))
    centre, vol, surf, fc, fa, nb = med.fv_geometry(xyz, cells)
    # a logically Cartesian n^3 mesh: 3 n^2 (n - 1) interior faces, 6 n^2 border faces
    assert len(cells) == ncells and len(fc) == nfaces and nb == nborder
    assert abs(vol.sum() - 1.0) < 1e-13 and vol.min() > 0
    fix = MS.load_fixture(name)
    assert np.array_equal(fc, fix["face_cells"])
    for got, key in ((centre, "centre"), (vol, "volume"), (surf, "surface"), (fa, "face_area")):
        assert np.allclose(got, fix[key], rtol=1e-14, atol=1e-16)
    # the Kershaw distortion: cell volumes spread over a factor > 30, faces are not planar
    assert vol.max() / vol.min() > 30
    # the one-call loader of the harness gives the same dictionary
    direct = MS.load_med(os.path.join(REF, rel))
    assert all(np.allclose(direct[k], fix[k], rtol=1e-14, atol=1e-16) for k in direct)


@needs_ref
def test_boundary_groups_of_the_reference_meshes():
    """Family numbers and group names (what Face::getGroupName() returns in the reference's assemblies)."""
    fam, per_geo = med.read_med_families(os.path.join(REF, "3DKershaw", "Kershaw1.med"))
    assert fam[-2] == ["boundary"] and fam[0] == []
    faces = per_geo["POG"]                               # 3 * 8 * 8 * 9 quadrilateral faces, the 6 * 8 * 8 border ones in "boundary"
    assert len(faces) == 1728 and int((faces == -2).sum()) == 384 and int((faces == -1).sum()) == 1344
    assert np.all(per_geo["POE"] == 0)
    fam, per_geo = med.read_med_families(os.path.join(REF, "meshCube.med"))
    sides = {names[0] for num, names in fam.items() if names}
    assert sides == {"Bas", "Haut", "Gauche", "Droite", "Devant", "Derriere"}
    assert len(per_geo["TR3"]) == 84 and set(np.unique(per_geo["TR3"])) <= set(fam)
