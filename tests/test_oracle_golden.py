"""The oracle against its own committed vectors (tests/golden/oracle_states.npz): a change in the restatement of the
reference's forms, in the quadrature, the lattice or the loop of ``run()`` shows up here, on the CPU, before any GPU
comparison.  The C/OpenMP restatement (second implementation, the timed CPU baseline) must land on the same vectors."""
import numpy as np
import pytest

from tests.golden_cases import CASES, OUTPUTS, load_case, run_oracle
from tests.util import rel_err


@pytest.mark.parametrize("dim,p", CASES)
def test_numpy_oracle_reproduces_golden(dim, p):
    c = load_case(dim, p)
    out = run_oracle(c)
    for k in OUTPUTS:
        assert rel_err(out[k], c[k]) < 1e-13, k
    assert np.array_equal(c["s_end"], np.swapaxes(c["s_end"], 2, 3))       # symmetric data stays symmetric


@pytest.mark.parametrize("dim,p", CASES)
def test_c_oracle_lands_on_golden(dim, p):
    from oracle.c_oracle import COracle
    from oracle.elastic_oracle import ElasticOracle
    c = load_case(dim, p)
    orc = ElasticOracle(c["coords"], c["cells"], int(c["degree"]), sigma_degree=int(c["sigma_degree"]))
    orc.l, orc.mu, orc.density, orc.dt = c["lam"], c["mu"], 1.0, float(c["dt"])
    orc.sigma = c["sigma"]
    co = COracle(orc)
    E, nd, d = orc.E, orc.nd, dim
    u, s = c["u0"].copy(), c["s0"].copy()
    for n in range(len(c["amp"])):
        src = np.zeros(E * nd * d * d)
        src[c["sdof"]] = c["amp"][n]
        co.step_inplace(u, s, src.reshape(E, nd, d, d), float(c["dt"]))
    assert rel_err(u, c["u_end"]) < 1e-12 and rel_err(s, c["s_end"]) < 1e-12
