"""The CUDA path against the committed golden vectors (tests/golden/oracle_states.npz), through the C ABI: per pass
<= 1e-12 relative L2, after the stored number of steps <= 1e-10 (BASELINE.json's tolerance), with full and with
symmetric stress storage.  Unlike tests/test_gpu_parity.py nothing of oracle/ runs here: the expected values are the
stored ones (pinned on the CPU by tests/test_oracle_golden.py)."""
import numpy as np
import pytest

from tests.golden_cases import CASES, load_case
from tests.util import rel_err

pytestmark = pytest.mark.gpu


def _device(c, dim, packed):
    from seigen_b200.device import DeviceSolver
    from seigen_b200.mesh import Mesh
    mesh = Mesh(np.ascontiguousarray(c["coords"]), np.ascontiguousarray(c["cells"]), name="golden")
    dev = DeviceSolver(mesh, int(c["degree"]), symmetric=packed)
    dev.set_material(1.0, c["lam"], c["mu"])
    dev.set_absorption(c["sigma"], int(c["sigma_degree"]))
    dev.set_source(c["sdof"], c["amp"])
    dev.set_state(c["u0"].reshape(-1, dim), c["s0"].reshape(-1, dim, dim))
    return dev


@pytest.mark.parametrize("dim,p", CASES)
@pytest.mark.parametrize("packed", [False, True])
def test_passes_and_steps_match_golden(dim, p, packed):
    from seigen_b200 import capi
    c = load_case(dim, p)
    dt = float(c["dt"])
    dev = _device(c, dim, packed)
    order = [(1, capi.FIELD_UH, "uh1"), (2, capi.FIELD_SH, "stemp"), (3, capi.FIELD_U, "u1"),
             (4, capi.FIELD_SH, "sh1"), (5, capi.FIELD_UH, "utemp"), (6, capi.FIELD_S, "s1")]
    for stage, field, key in order:
        dev.stage(stage, dt, 0)
        assert rel_err(dev.get_field(field).reshape(c[key].shape), c[key]) < 1e-12, key
    dev.set_state(c["u0"].reshape(-1, dim), c["s0"].reshape(-1, dim, dim))
    dev.step(len(c["amp"]), dt, 0)
    u, s = dev.get_state()
    assert rel_err(u.reshape(c["u_end"].shape), c["u_end"]) < 1e-10
    assert rel_err(s.reshape(c["s_end"].shape), c["s_end"]) < 1e-10
    dev.close()
