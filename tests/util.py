"""Shared helpers for the test-suite (test infrastructure; may import ``oracle``)."""
import numpy as np

from oracle.elastic_oracle import ElasticOracle
from oracle.nodal import NodalOperator
from seigen_b200.mesh import BoxMesh, IntervalMesh, RectangleMesh, perturb_vertices
from seigen_b200.refelem import get_refelem


def small_mesh(dim, n=None, perturb=0.15, seed=0):
    if dim == 1:
        m = IntervalMesh(n or 37, 1.7)
    elif dim == 2:
        n = n or 4
        m = RectangleMesh(n, n + 1, 1.3, 1.0)
    else:
        n = n or 2
        m = BoxMesh(n, n + 1, n, 1.0, 1.2, 0.9)
    return perturb_vertices(m, perturb, seed) if perturb else m


def nodal_from_mesh(mesh, degree):
    el = get_refelem(mesh.dim, degree)
    t = mesh.topology
    return NodalOperator(el.Dr, el.Lift, el.fnodes, el.ftab, t.nbr, t.code, t.jinv), el


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def random_state(mesh, degree, seed=1):
    el = get_refelem(mesh.dim, degree)
    rng = np.random.default_rng(seed)
    E, d = mesh.num_cells(), mesh.dim
    return rng.standard_normal((E, el.nd, d)), rng.standard_normal((E, el.nd, d, d))
