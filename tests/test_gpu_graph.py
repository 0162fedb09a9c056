"""A whole step -- advection! + move_particles! + inject_particles! + particle2grid! (+ phase_ratios_center!), both hand-offs on --
captured into a CUDA graph (include/justpic_c.h, JP_OPT_GRAPH_STEP_OFFSET; api.capture_step) and replayed: every replay must leave
exactly what the eager calls leave, which the oracle checks step by step.  What cannot be captured is refused with a message."""
import numpy as np
import pytest
import torch

from tests.problems import stream_velocity, cfl_dt, vertex_field_linear
from tests.test_gpu_parity import Twin, jp, dev, host, assert_same, assert_close, ids

pytestmark = pytest.mark.gpu


def _problem(g, K=2):
    J = jp()
    t = Twin(*g, nxcell=12, max_xcell=24, min_xcell=8)
    V = stream_velocity(t.gr); Vd = [dev(v) for v in V]
    dt = cfl_dt(t.gr, V, 0.9)
    T = vertex_field_linear(t.gr)
    pT, ph = J.init_cell_arrays(t.p, 2)
    J.grid2particle(pT, dev(T), t.p)
    opT = np.zeros_like(t.co[0]); t.o.grid2particle(t.co, t.idx, opT, T)
    oph = np.where(t.idx > 0, 1.0 + (t.co[0] < t.co[-1]), 0.0)
    ph.copy_(dev(oph))
    F = dev(np.zeros_like(T))
    pr = J.PhaseRatios(J.CUDABackend, K, t.gr.n)
    J.move_interp_handoff(t.p, Fp=pT, phases=ph, nphases=K)

    def step():
        J.advection(t.p, J.RungeKutta2(), Vd, dt, classify=True)
        J.move_particles(t.p, (pT, ph))
        J.inject_particles(t.p, (pT, ph))
        J.particle2grid(F, pT, t.p)
        J.phase_ratios_center(pr, t.p, ph)

    def oracle_step(it):
        t.o.advect(t.co, t.idx, 1, 0.5, V, dt)
        st = t.o.move(t.co, t.idx, [opT, oph])
        inj = t.o.inject(t.co, t.idx, [opT, oph], t.min_xcell, t.seed, it)
        oF = np.empty_like(T); t.o.particle2grid(t.co, t.idx, oF, opT)
        ratios = np.zeros(t.o.cell_shape(K)); t.o.phase_ratios_center(t.co, ratios, oph, K)
        return st, inj, oF, ratios

    return J, t, step, oracle_step, (pT, ph), (opT, oph), F, pr


@pytest.mark.parametrize("g", [(2, 24, True), (2, (19, 33), False), (3, 10, True), (3, (9, 7, 12), False)], ids=ids)
def test_captured_step_replays_like_the_eager_calls(g):
    J, t, step, oracle_step, gargs, oargs, F, pr = _problem(g)
    graph = J.capture_step(t.p, step, warmup=2)            # eager steps 0, 1; the captured call is step 2 (not executed)
    for it in range(2):
        oracle_step(it)
    for it in range(2, 8):
        graph.replay()
        st, inj, oF, ratios = oracle_step(it)
        t.check_state(f"replay {it - 2} (step {it})", gargs, oargs)
        assert J.move_stats(t.p) == st and J.inject_stats(t.p) == inj
        assert J.last_move_path(t.p) == "plan" and J.last_move_classify(t.p) == "handoff"
        assert_close(F, oF, f"step {it}: particle2grid inside the graph")
        assert_same(pr.center, ratios, f"step {it}: phase_ratios_center inside the graph")
    assert J.graph_step_offset(t.p) == 6
    # back to eager calls: the device-side step offset is reset, the host-side counter goes on where the replays stopped
    J.graph_step_offset(t.p, 0)
    t.p._inject_step = 8
    step(); oracle_step(8)
    t.check_state("eager step after the replays", gargs, oargs)


def test_capture_refuses_what_it_cannot_do():
    J = jp()
    t = Twin(2, 24, True)
    pT, = J.init_cell_arrays(t.p, 1)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        g.capture_begin()
        try:
            with pytest.raises(ValueError, match="eagerly"):
                J.move_particles(t.p, (pT,))               # the plan workspace does not exist yet
        finally:
            g.capture_end()
    torch.cuda.current_stream().wait_stream(s)
    J.move_particles(t.p, (pT,))                           # ... and the context is intact
    torch.cuda.synchronize()
    assert J.last_move_path(t.p) == "plan"
