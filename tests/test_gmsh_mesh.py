"""Unstructured meshes: Gmsh reader + the general adjacency path on Delaunay meshes with mixed cell
orientation, checked with the nodal CPU operator against the literal oracle (CPU) and on the GPU (gpu mark)."""
import numpy as np
import pytest
from scipy.spatial import Delaunay

from oracle.elastic_oracle import ElasticOracle
from seigen_b200.mesh import Mesh, read_gmsh, write_gmsh
from tests.util import nodal_from_mesh, random_state, rel_err


def delaunay_mesh(dim, n=60, seed=0):
    rng = np.random.default_rng(seed)
    pts = rng.uniform(size=(n, dim)) * ([300.0, 150.0] if dim == 2 else [1.0, 1.0, 1.0])
    corners = np.array(np.meshgrid(*[[0.0, 1.0]] * dim)).reshape(dim, -1).T * pts.max(axis=0)
    pts = np.vstack([pts, corners])
    tri = Delaunay(pts)
    cells = tri.simplices.copy()
    vol = np.abs(np.linalg.det(pts[cells[:, 1:]] - pts[cells[:, :1]]))
    cells = cells[vol > 1e-6 * vol.max()]                   # drop slivers on the hull
    cells[::3, [0, 1]] = cells[::3, [1, 0]]                 # mixed orientation
    return pts, cells.astype(np.int32)


@pytest.mark.parametrize("dim", [2, 3])
def test_gmsh_roundtrip_and_operator(tmp_path, dim):
    pts, cells = delaunay_mesh(dim)
    path = tmp_path / "domain.msh"
    write_gmsh(path, pts, cells)
    mesh = Mesh(str(path))
    assert mesh.dim == dim and mesh.num_cells() == len(cells)
    assert np.allclose(mesh.coords[mesh.cells], pts[cells])
    orc = ElasticOracle(mesh.coords, mesh.cells, 2)
    orc.l, orc.mu = 0.7, 0.3
    op, _ = nodal_from_mesh(mesh, 2)
    u, s = random_state(mesh, 2)
    assert rel_err(op.Dv(s), orc.solve_f(s, u)) < 1e-11
    assert rel_err(op.Ds(u, 0.7, 0.3), orc.solve_g(u, None)) < 1e-11


def test_gmsh_v4_and_lower_dimensional_elements(tmp_path):
    path = tmp_path / "tiny.msh"
    path.write_text("""$MeshFormat
4.1 0 8
$EndMeshFormat
$Nodes
1 4 1 4
2 1 0 4
1
2
3
4
0 0 0
1 0 0
1 1 0
0 1 0
$EndNodes
$Elements
2 3 1 3
1 1 1 1
1 1 2
2 1 2 2
2 1 2 3
3 1 3 4
$EndElements
""")
    coords, cells = read_gmsh(path)
    assert coords.shape == (4, 2) and cells.tolist() == [[0, 1, 2], [0, 2, 3]]


@pytest.mark.gpu
@pytest.mark.parametrize("dim,p", [(2, 2), (3, 1)])
def test_unstructured_gpu_parity(dim, p):
    from seigen_b200 import capi
    from seigen_b200.device import DeviceSolver
    pts, cells = delaunay_mesh(dim, n=400 if dim == 2 else 120, seed=3)
    mesh = Mesh(pts, cells)
    orc = ElasticOracle(mesh.coords, mesh.cells, p)
    orc.l, orc.mu, orc.density, orc.dt = 0.5, 0.25, 1.0, 1e-4 * (300 if dim == 2 else 1)
    dev = DeviceSolver(mesh, p)
    dev.set_material(1.0, 0.5, 0.25)
    u0, s0 = random_state(mesh, p)
    dev.set_state(u0.reshape(-1, dim), s0.reshape(-1, dim, dim))
    u, s = u0, s0
    for _ in range(3):
        u, s, _ = orc.step(u, s, 0.0)
    dev.step(3, orc.dt, 0)
    ug, sg = dev.get_state()
    assert rel_err(ug.reshape(u.shape), u) < 1e-11
    assert rel_err(sg.reshape(s.shape), s) < 1e-11
    dev.close()
