"""Pins the CPU oracle to the reference's only stored solution data for this path: the external sensor trace
tests/explosive_source/REF-C1 (fixture tests/golden/ref_c1.npz, made by scripts/make_golden.py), which
tests/explosive_source/uy.py:36-43 overlays on -u_y at (45, 149).  The scenario of
tests/explosive_source/explosive_source_lf4.py runs on a 100 x 50 m sub-domain below the same free surface with the
stable time step.  REF-C1 comes from a different solver: measured agreement 12.6 % relative
L2 over 0.1 <= t <= 0.45, peak ratio 1.047, peak 3.6 ms late -- the tolerances below are those numbers plus a
margin; the reference itself only checks by eye.  The full 300 x 150 domain to T = 2.5 s (all three sensors) takes
minutes on the CPU, so its traces are stored (tests/golden/oracle_refc_traces.npz, scripts/make_golden_traces.py)
and compared with REF-C1..3 here; the GPU reproduces those traces in tests/test_gpu_fullsize.py."""
import os

import numpy as np

from oracle.c_oracle import COracle
from oracle.elastic_oracle import step_times
from tests.scenarios import explosive_oracle, locate

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_c1.npz")


def sensor_trace_oracle(T=0.45, Lx=100.0, Ly=50.0, h=2.5):
    mesh, orc, src = explosive_oracle(Lx, Ly, h)
    co = COracle(orc)
    e, xi = locate(mesh.coords, mesh.cells, (45.0, Ly - 1.0))
    phi = orc.el.tab(xi[None])[0]
    u = np.zeros((orc.E, orc.nd, 2))
    s = np.zeros((orc.E, orc.nd, 2, 2))
    times = step_times(T, orc.dt)
    trace = []
    for t in times:
        co.step_inplace(u, s, src(t), orc.dt)
        trace.append(-(phi @ u[e, :, 1]))
    return np.array(times), np.array(trace)


def compare_with_ref(times, trace):
    z = np.load(GOLDEN)
    ref = np.interp(times, z["t"], z["uy"])
    w = (times >= 0.1) & (times <= 0.45)
    rel = np.linalg.norm(trace[w] - ref[w]) / np.linalg.norm(ref[w])
    peak_sim, peak_ref = np.abs(trace[w]).max(), np.abs(ref[w]).max()
    t_sim, t_ref = times[w][np.abs(trace[w]).argmax()], times[w][np.abs(ref[w]).argmax()]
    return rel, peak_sim / peak_ref, t_sim - t_ref


def test_oracle_tracks_ref_c1():
    times, trace = sensor_trace_oracle()
    rel, peak_ratio, dt_peak = compare_with_ref(times, trace)
    assert np.isfinite(trace).all() and np.abs(trace).max() < 1e-3          # stable (Courant 0.5 blows up, App. B-7)
    assert rel < 0.15, rel
    assert 0.95 < peak_ratio < 1.15, peak_ratio
    assert abs(dt_peak) < 0.006, dt_peak


def test_stored_full_domain_traces_track_ref_c123():
    """The oracle's traces on the shipped 120 x 60 mesh to T = 2.5 s, 0.3 m beside the sensors (interior points):
    waveform correlation with REF-C1..3 >= 0.97 in the windows uy.py plots; amplitude 0.86 x at C1 (the near field
    varies quickly beside the source; at the sensor itself it is 1.05 x, see test_oracle_tracks_ref_c1) and 2.2 x at
    C2 / C3 (Rayleigh wave from a source 1 m deep on an h = 2.5 m mesh: not resolved -- documented, not hidden)."""
    zt = np.load(os.path.join(os.path.dirname(GOLDEN), "oracle_refc_traces.npz"))
    zr = np.load(os.path.join(os.path.dirname(GOLDEN), "ref_c123.npz"))
    tt = zt["t"]
    assert len(tt) == 2078
    expected_amp = [(0.8, 0.95), (2.0, 2.45), (2.0, 2.45)]
    for k, (lo, hi) in enumerate([(0.1, 1.0), (0.5, 1.5), (1.0, 2.5)]):
        ref = np.interp(tt, zr["t"], zr["uy"][k])
        w = (tt >= lo) & (tt <= hi)
        sim = -zt["u"][w, k, 1]
        corr = sim @ ref[w] / (np.linalg.norm(sim) * np.linalg.norm(ref[w]))
        amp = (sim @ ref[w]) / (ref[w] @ ref[w])
        assert corr > 0.97, (k, corr)
        assert expected_amp[k][0] < amp < expected_amp[k][1], (k, amp)
