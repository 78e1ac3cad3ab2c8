"""The replica-exchange driver end to end on one GPU: a 6-state schedule, sample lines in the reference's .out format
(ref: README.md:193-201, example/abfe/abfe.py:149-160), State-XML checkpoints (ref: example/abfe/temoa-g1-equil.xml) and a
restart that continues exactly where the first run stopped."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_driver_writes_samples_checkpoints_and_restarts(tmp_path):
    import atmmetaforce as atm
    from atmmetaforce import synthetic, io
    s = synthetic.water_box(6000, n_lig=15, seed=6)
    sched = synthetic.atm_schedule_22()[3:9]

    def make(outdir):
        return atm.ReplicaExchangeDriver(s, sched, "job", outdir=str(outdir), temperature=300.0, steps_per_cycle=12, prune_every=5,
                                         rebuild_every=10, skin=0.1, skin_outer=0.1, checkpoint_every=2, seed=5)

    # reference run: 4 cycles in one go
    a = make(tmp_path / "a")
    sa = a.run(4)
    state_a, cycle_a, step_a = a.rex.state_dict(), a.cycle, a.step_no
    a.close()
    # the same, interrupted after 2 cycles and restarted from the checkpoint
    b = make(tmp_path / "b")
    sb = b.run(2)
    b.close()
    c = make(tmp_path / "b")
    assert c.restart() and c.cycle == 2 and c.step_no == 24
    sc = c.run(2)
    assert (c.cycle, c.step_no) == (cycle_a, step_a) == (4, 48)
    assert c.rex.state_dict()["replica_state"] == state_a["replica_state"]
    c.close()
    # sample lines: 9 columns, 4 per replica, the lambda columns are those of the state the replica held
    for g in range(6):
        lines = open(tmp_path / "a" / f"job_r{g}.out").read().splitlines()
        assert len(lines) == 4
        lines_b = open(tmp_path / "b" / f"job_r{g}.out").read().splitlines()
        assert len(lines_b) == 4
        for ln, (cyc, gg, st, pe, u) in zip(lines, [x for x in sa if x[1] == g]):
            rec = io.parse_sample_line(ln)
            assert rec["temperature"] == 300.0 and abs(rec["lambda1"] - sched[st][0]) < 1e-6 and abs(rec["lambda2"] - sched[st][1]) < 1e-6
            assert abs(rec["pert_energy"] - u) <= 1e-4 * max(1.0, abs(u)) and np.isfinite(rec["pot_energy"])
        # the propagator is seeded: the restarted trajectory reproduces the uninterrupted one, sample for sample
        assert lines == lines_b
    assert sorted(state_a["replica_state"]) == list(range(6)) and state_a["cycle"] == 4
    # checkpoints: State XML with the nine ATM parameters of the state each replica holds, plus the exchange bookkeeping
    meta = json.load(open(tmp_path / "a" / "job-hrex.json"))
    assert meta["cycle"] == 4 and meta["step"] == 48 and meta["hrex"]["replica_state"] == state_a["replica_state"]
    st = io.read_state_xml(tmp_path / "a" / "job_r2-chk.xml")
    row = sched[state_a["replica_state"][2]]
    assert st["positions"].shape == (s["pos"].shape[0], 3) and abs(st["parameters"]["ATMLambda2"] - row[1]) < 1e-12
    assert set(st["parameters"]) == {"ATMLambda1", "ATMLambda2", "ATMAlpha", "ATMU0", "ATMW0", "ATMUmax", "ATMUbcore", "ATMAcore", "ATMDirection"}
