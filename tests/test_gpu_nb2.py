"""GPU parity of the fused two-state direct-space path (atm_step) against the CPU oracle.

Bars (BASELINE.json north_star): energies <= 1e-6 relative, per-atom forces <= 1e-5 relative RMS, against a
double-precision evaluation of the SAME float-rounded coordinates.  The tolerances asserted here are about twice the
largest error measured on B200 (profiles/r2_parity_errors.md: U1 8.5e-8 relative, forces 3.5e-6 relative RMS, sp 5.8e-7,
u 3.1e-4 kJ/mol on the reference fixtures and 1.5e-3 on the synthetic 23k system), i.e. tighter than the bars.  The reference's only pinned number on this path is
u = 58.2 +- 0.1 kJ/mol for the TEMOA-G1 fixture (python/tests/test_abfe.py:148-150).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

E_U1, E_U2, E_U, E_USC, E_EBIAS, E_ENERGY, E_SP, E_NPAIRS = range(8)


def _run(s, cutoff, alpha, params, skin=0.1, perm=None, du_ext=0.0, with_ext=False):
    import torch
    import atmmetaforce as atm
    import oracle_py as O
    from helpers import oracle_system, make_backend, force_from_fixed
    n = s["pos"].shape[0]
    be, posq, atom_index = make_backend(atm, s, cutoff, alpha, params, skin=skin, perm=perm)
    P = be.P
    be.rebuild(posq)
    force = torch.zeros((1, 3 * P), dtype=torch.int64, device="cuda")
    ext = torch.tensor([[0.0, du_ext]], dtype=torch.float64, device="cuda") if with_ext else None
    be.step(posq, force, energy_ext=ext)
    en = be.get_energies()[0]
    f_slot = force_from_fixed(force.cpu().numpy()[0], n, P)
    f_gpu = np.zeros_like(f_slot)
    f_gpu[atom_index] = f_slot
    stats = be.nb_stats()
    # second step must give bit-identical forces (accumulators are re-zeroed, fixed point is order independent)
    force2 = torch.zeros_like(force)
    be.step(posq, force2, energy_ext=ext)
    torch.cuda.synchronize()
    assert torch.equal(force, force2)
    en2 = be.get_energies()[0]
    assert np.array_equal(en[:7], en2[:7])
    be.close()
    # oracle on the float-rounded inputs the GPU saw
    pos32 = s["pos"].astype(np.float32).astype(np.float64)
    d32 = s["displ"].astype(np.float32)
    pos2_32 = (s["pos"].astype(np.float32) + d32).astype(np.float64)  # the float add of CopyState
    S = oracle_system(O, s, cutoff, alpha)
    e1, c1, f1 = S.nb_direct(pos32)
    e2, c2, f2 = S.nb_direct(pos2_32)
    sc = O.scalars(params, e1, e2 + du_ext)
    f_ref = O.merge_ref(np.zeros_like(f1), f1, f2, sc["sp_ref"], params[8])
    return dict(en=en, f_gpu=f_gpu, f_ref=f_ref, e1=e1, e2=e2 + du_ext, sc=sc, stats=stats, f1=f1, f2=f2)


def _check(r, tol_u=1e-3):
    from helpers import rel_rms
    en, sc = r["en"], r["sc"]
    assert abs(en[E_U1] - r["e1"]) <= 2e-7 * abs(r["e1"]), (en[E_U1], r["e1"])
    assert abs(en[E_U2] - r["e2"]) <= 2e-7 * abs(r["e2"])
    du = r["e2"] - r["e1"]
    assert abs((en[E_U2] - en[E_U1]) - du) <= max(tol_u, 2e-7 * abs(r["e1"]))
    assert abs(en[E_USC] - sc["u_sc"]) <= tol_u
    assert abs(en[E_SP] - sc["sp"]) <= 2e-6
    assert abs(en[E_ENERGY] - sc["energy"]) <= 2e-7 * abs(sc["energy"]) + tol_u
    err = rel_rms(r["f_gpu"], r["f_ref"])
    from helpers import record
    record(U1_rel=abs(en[E_U1] - r["e1"]) / abs(r["e1"]), U2_rel=abs(en[E_U2] - r["e2"]) / abs(r["e2"]),
           u_abs=abs((en[E_U2] - en[E_U1]) - du), usc_abs=abs(en[E_USC] - sc["u_sc"]), sp_abs=abs(en[E_SP] - sc["sp"]), force_rel_rms=err)
    assert err <= 7e-6, err
    return err


def test_abfe_fixture_pin(abfe):
    """u(direct) + oracle reciprocal difference reproduces the reference pin 58.2 +- 0.1 kJ/mol."""
    import oracle_py as O
    from helpers import oracle_system
    alpha = O.ewald_alpha(1.0)
    S = oracle_system(O, abfe, 1.0, alpha)
    r1, _ = S.ewald_recip(abfe["pos"], 1e-10)
    r2, _ = S.ewald_recip(abfe["pos"] + abfe["displ"], 1e-10)
    res = _run(abfe, 1.0, alpha, abfe["params"], du_ext=r2 - r1, with_ext=True)
    err = _check(res)
    u = res["en"][E_USC]
    print("abfe: u_sc = %.4f (pin 58.2), U1 = %.3f, force rel rms = %.2e, stats %s" % (u, res["en"][E_U1], err, res["stats"]))
    assert abs(u - 58.2) <= 0.1
    # direct-space components quoted in SURVEY.md section 7.5 T5
    assert abs((res["en"][E_U2] - res["en"][E_U1]) - (r2 - r1) - 71.8773) <= 0.01


def test_rbfe_fixture(rbfe):
    import oracle_py as O
    alpha = O.ewald_alpha(1.0)
    res = _run(rbfe, 1.0, alpha, rbfe["params"])
    err = _check(res)
    print("rbfe: u = %.4f, force rel rms = %.2e, stats %s" % (res["en"][E_U], err, res["stats"]))
    assert res["stats"]["groups"] == 2 and res["stats"]["displaced_atoms"] == 38


@pytest.mark.parametrize("params", [
    [0.2, 0.7, 0.02, 40.0, 3.0, 300.0, 20.0, 0.0625, 1.0],    # alpha > 0, soft-core branch active (ubcore < u)
    [0.2, 0.7, 0.02, 40.0, 3.0, 300.0, 20.0, 0.0625, -1.0],   # direction -1
    [0.0, 0.0, 0.0, 0.0, 0.0, 800.0, 400.0, 0.0625, 1.0],     # lambda = 0: merged force == F1
    [1.0, 1.0, 0.0, 0.0, 0.0, 1e9, 5e8, 0.0625, 1.0],         # lambda = 1, no soft core: merged force == F2
])
def test_abfe_parameter_branches(abfe, params):
    import oracle_py as O
    from helpers import rel_rms
    alpha = O.ewald_alpha(1.0)
    res = _run(abfe, 1.0, alpha, params)
    _check(res)
    if params[0] == 0.0 and params[1] == 0.0:
        assert rel_rms(res["f_gpu"], res["f1"]) <= 1e-5
    if params[0] == 1.0:
        assert rel_rms(res["f_gpu"], res["f2"]) <= 1e-5


def test_abfe_reordered_slots(abfe):
    """OpenMM reorders atoms; slot s holds atom atom_index[s].  Results must not depend on the slot order."""
    import oracle_py as O
    rng = np.random.default_rng(11)
    perm = rng.permutation(abfe["pos"].shape[0]).astype(np.int32)
    res = _run(abfe, 1.0, O.ewald_alpha(1.0), abfe["params"], perm=perm)
    _check(res)


def test_no_displacement_is_single_state():
    """Zero displacement: U2 == U1 exactly, u == 0, merged force == F1."""
    import oracle_py as O
    from atmmetaforce import synthetic
    s = synthetic.water_box(6000, n_lig=0)
    params = [0.5, 0.5, 0, 0, 0, 800, 400, 0.0625, 1.0]
    res = _run(s, s["cutoff"], s["ewald_alpha"], params)
    _check(res)
    assert res["en"][E_U] == 0.0 and res["en"][E_U1] == res["en"][E_U2]


@pytest.mark.parametrize("maker", ["config3", "water20k"])
def test_synthetic_systems(maker):
    from atmmetaforce import synthetic
    s = synthetic.config3() if maker == "config3" else synthetic.water_box(20000)
    params = synthetic.atm_schedule_22()[7]
    res = _run(s, s["cutoff"], s["ewald_alpha"], params, skin=0.1)
    err = _check(res, tol_u=5e-3)
    print(maker, "u = %.4f, force rel rms = %.2e, stats %s" % (res["en"][E_U], err, res["stats"]))


def test_small_box_is_rejected():
    import atmmetaforce as atm
    from atmmetaforce import synthetic
    from helpers import make_backend
    s = synthetic.water_box(1500, n_lig=0)   # ~2.4 nm box < 2*(0.9+0.1)+cluster extent
    with pytest.raises(atm.ATMError):
        be, posq, _ = make_backend(atm, s, 1.2, 2.0, [0.5, 0.5, 0, 0, 0, 800, 400, 0.0625, 1.0])
        be.rebuild(posq)


def test_dual_list_prune_after_motion(abfe):
    """Outer list with a wide skin, atoms then move by < skin/2 and only the cheap prune runs: results must match an
    oracle evaluation at the NEW coordinates (validates list reuse, the per-step image shift and the prune)."""
    import torch
    import atmmetaforce as atm
    import oracle_py as O
    from helpers import oracle_system, make_backend, force_from_fixed, rel_rms
    s = dict(abfe)
    alpha = O.ewald_alpha(1.0)
    params = abfe["params"]
    n = s["pos"].shape[0]
    import atmmetaforce.backend as B
    be = atm.ATMBackend(n, precision="mixed")
    be.set_displacements(s["displ"])
    be.set_box(s["box"])
    be.set_parameters(params)
    be.nb_setup(s["charge"], s["sigma"], s["epsilon"], 1.0, alpha, skin=0.05, skin_outer=0.3, exclusions=s["excl"],
                exception_pairs=s["exc14"], exception_params=s["exc14_par"])
    P = be.P
    posq = np.zeros((1, P, 4), np.float32)
    posq[0, :n, :3] = s["pos"]
    posq[0, :n, 3] = s["charge"]
    d_posq = torch.from_numpy(posq).cuda()
    be.rebuild(d_posq)
    rng = np.random.default_rng(5)
    S = oracle_system(O, s, 1.0, alpha)
    moved = posq.copy()
    for it in range(3):
        # total drift since the rebuild stays below skin_outer/2 = 0.15; since the last prune below skin/2 = 0.025
        moved[0, :n, :3] += rng.uniform(-0.012, 0.012, (n, 3)).astype(np.float32)
        d_moved = torch.from_numpy(moved).cuda()
        be.prune(d_moved)
        force = torch.zeros((1, 3 * P), dtype=torch.int64, device="cuda")
        be.step(d_moved, force)
        en = be.get_energies()[0]
        pos32 = moved[0, :n, :3].astype(np.float64)
        pos2_32 = (moved[0, :n, :3] + s["displ"].astype(np.float32)).astype(np.float64)
        e1, _, f1 = S.nb_direct(pos32)
        e2, _, f2 = S.nb_direct(pos2_32)
        sc = O.scalars(params, e1, e2)
        f_ref = O.merge_ref(np.zeros_like(f1), f1, f2, sc["sp_ref"], params[8])
        f_gpu = force_from_fixed(force.cpu().numpy()[0], n, P)
        assert abs(en[E_U1] - e1) <= 1e-6 * abs(e1)
        assert abs(en[E_U] - (e2 - e1)) <= 5e-3
        assert rel_rms(f_gpu, f_ref) <= 1e-5
    be.close()


def test_batched_replicas_match_oracle_individually():
    """R = 3 replicas with different coordinates and different lambda states in ONE handle: each must match its own
    oracle evaluation (the batching is invisible)."""
    import torch
    import atmmetaforce as atm
    import oracle_py as O
    from atmmetaforce import synthetic
    from helpers import oracle_system, force_from_fixed, rel_rms
    s = synthetic.water_box(9000, n_lig=30, seed=21)
    n = s["pos"].shape[0]
    sched = synthetic.atm_schedule_22()
    states = [3, 9, 15]     # includes a direction = -1 state
    R = 3
    be = atm.ATMBackend(n, precision="mixed", num_replicas=R)
    P = be.P
    be.set_displacements(s["displ"])
    be.set_box(s["box"])
    for r in range(R):
        be.set_parameters(sched[states[r]], replica=r)
    be.nb_setup(s["charge"], s["sigma"], s["epsilon"], s["cutoff"], s["ewald_alpha"], skin=0.08, skin_outer=0.2, exclusions=s["excl"])
    rng = np.random.default_rng(8)
    posq = np.zeros((R, P, 4), np.float32)
    for r in range(R):
        posq[r, :n, :3] = s["pos"] + rng.normal(0, 0.01, (n, 3))
        posq[r, :n, 3] = s["charge"]
    d_posq = torch.from_numpy(posq).cuda()
    force = torch.zeros((R, 3 * P), dtype=torch.int64, device="cuda")
    be.rebuild(d_posq)
    be.step(d_posq, force)
    en = be.get_energies()
    S = oracle_system(O, s, s["cutoff"], s["ewald_alpha"])
    d32 = s["displ"].astype(np.float32)
    for r in range(R):
        p1 = posq[r, :n, :3].astype(np.float64)
        p2 = (posq[r, :n, :3] + d32).astype(np.float64)
        e1, _, f1 = S.nb_direct(p1)
        e2, _, f2 = S.nb_direct(p2)
        prm = sched[states[r]]
        sc = O.scalars(prm, e1, e2)
        f_ref = O.merge_ref(np.zeros_like(f1), f1, f2, sc["sp_ref"], prm[8])
        f_gpu = force_from_fixed(force.cpu().numpy()[r], n, P)
        assert abs(en[r, E_U1] - e1) <= 1e-6 * abs(e1)
        assert abs(en[r, E_USC] - sc["u_sc"]) <= 5e-3
        assert abs(en[r, E_SP] - sc["sp"]) <= 1e-4
        assert rel_rms(f_gpu, f_ref) <= 1e-5
    be.close()


def test_config4_100k_rbfe():
    """BASELINE configs[3]: ~100k atoms, two 40-atom ligands displaced +d / -d (two displacement groups)."""
    from atmmetaforce import synthetic
    s = synthetic.config4()
    assert s["pos"].shape[0] > 90000
    params = synthetic.atm_schedule_22()[5]
    res = _run(s, s["cutoff"], s["ewald_alpha"], params, skin=0.1)
    err = _check(res, tol_u=5e-3)
    assert res["stats"]["groups"] == 2 and res["stats"]["displaced_atoms"] == 80
    print("config4 u = %.4f, force rel rms = %.2e, stats %s" % (res["en"][E_U], err, res["stats"]))


def test_single_atom_ligand_and_graph_replay():
    """One displaced atom (a monatomic ion); the CUDA-graph replay of the step must equal the plain launches bit for bit."""
    import torch
    import atmmetaforce as atm
    from atmmetaforce import synthetic
    s = synthetic.water_box(7000, n_lig=0, seed=2)
    n = s["pos"].shape[0]
    s["displ"][0] = [1.7, -1.1, 0.9]          # displace the oxygen of the first water only (its exclusions split states)
    params = [0.3, 0.6, 0.02, 20.0, 0.0, 800.0, 400.0, 0.0625, 1.0]
    res = _run(s, s["cutoff"], s["ewald_alpha"], params)
    _check(res, tol_u=5e-3)
    assert res["stats"]["displaced_atoms"] == 1
    from helpers import make_backend
    be, posq, _ = make_backend(atm, s, s["cutoff"], s["ewald_alpha"], params)
    be.rebuild(posq)
    f_plain = torch.zeros((1, 3 * be.P), dtype=torch.int64, device="cuda")
    f_graph = torch.zeros_like(f_plain)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        be.step(posq, f_plain, stream=st)
        for _ in range(3):
            f_graph.zero_()
            be.step(posq, f_graph, graph=True, stream=st)
    st.synchronize()
    assert torch.equal(f_plain, f_graph)
    be.close()


def test_large_displaced_group():
    """A displaced group that spans the box (every third water of a slab, ~1500 atoms): ligand classes are binned into
    xy columns like the environment, so no cluster outgrows the half-box limit."""
    from atmmetaforce import synthetic
    s = synthetic.water_box(12000, n_lig=0, seed=9)
    n = s["pos"].shape[0]
    pos = s["pos"]
    sel = np.where((pos[:, 2] < 0.35 * s["box"][2]))[0]
    sel = sel[(sel // 3) % 2 == 0]                     # whole molecules (O,H,H share index // 3)
    # the water generator is a jittered cubic lattice: shift by (k + 1/2) lattice spacings so that displaced molecules
    # land in interstitial positions instead of on top of other molecules (overlaps would overflow any fixed-point force)
    nl = int(round((n / 3) ** (1.0 / 3.0)))
    a = s["box"][0] / nl
    s["displ"][sel] = [0.5 * a, 0.5 * a, (int(0.45 * nl) + 0.5) * a]
    assert 1000 < sel.size < 4000
    params = [0.4, 0.4, 0.0, 0.0, 0.0, 1e7, 5e6, 0.0625, 1.0]   # the overlap energy is huge: keep the soft core out
    res = _run(s, s["cutoff"], s["ewald_alpha"], params, skin=0.05)
    from helpers import rel_rms
    en = res["en"]
    print("large group: U1 %.2f/%.2f U2 %.2f/%.2f frms %.2e stats %s" % (en[E_U1], res["e1"], en[E_U2], res["e2"],
          rel_rms(res["f_gpu"], res["f_ref"]), res["stats"]))
    assert abs(en[E_U1] - res["e1"]) <= 1e-6 * abs(res["e1"])
    assert abs(en[E_U2] - res["e2"]) <= 1e-6 * abs(res["e2"])
    assert rel_rms(res["f_gpu"], res["f_ref"]) <= 1e-5
    assert res["stats"]["displaced_atoms"] == sel.size


def test_unwrapped_coordinates_and_per_replica_boxes():
    """OpenMM keeps molecules whole and re-wraps them by their centres, so atoms sit up to about one box length outside
    the primary cell: whole molecules translated by -1/0/+1 box vectors must give the oracle's minimum-image result.
    (Farther out the fp32 coordinates themselves lose the digits the 1e-5 force bar needs -- at +-3 box lengths the
    measured deviation is 1.2e-5 -- which is why OpenMM re-wraps.)  Replica 1 also lives in its own (2 % larger) box, as under a
    barostat (the box mirror of copyState, CommonATMMetaForceKernels.cpp:214-219, is per context)."""
    import torch
    import atmmetaforce as atm
    import oracle_py as O
    from atmmetaforce import synthetic
    from helpers import force_from_fixed, rel_rms
    s = synthetic.water_box(9000, n_lig=30, seed=33)
    n = s["pos"].shape[0]
    sched = synthetic.atm_schedule_22()
    rng = np.random.default_rng(5)
    # molecule id per atom: the ligand (displaced atoms) is one molecule, every water (3 consecutive atoms) another
    lig = np.nonzero(np.abs(s["displ"]).sum(1) > 0)[0]
    mol = np.full(n, -1)
    mol[lig] = 0
    rest = np.nonzero(mol < 0)[0]
    mol[rest] = 1 + np.arange(rest.size) // 3
    shifts = rng.integers(-1, 2, (mol.max() + 1, 3)).astype(np.float64)
    scale = [1.0, 1.02]
    R = 2
    be = atm.ATMBackend(n, precision="mixed", num_replicas=R)
    P = be.P
    be.set_displacements(s["displ"])
    for r in range(R):
        be.set_box(s["box"] * scale[r], replica=r)
        be.set_parameters(sched[4 + 11 * r], replica=r)
    be.nb_setup(s["charge"], s["sigma"], s["epsilon"], s["cutoff"], s["ewald_alpha"], skin=0.1, exclusions=s["excl"])
    posq = np.zeros((R, P, 4), np.float32)
    for r in range(R):
        box = s["box"] * scale[r]
        posq[r, :n, :3] = s["pos"] * scale[r] + shifts[mol] * box
        posq[r, :n, 3] = s["charge"]
    d_posq = torch.from_numpy(posq).cuda()
    force = torch.zeros((R, 3 * P), dtype=torch.int64, device="cuda")
    be.rebuild(d_posq)
    be.step(d_posq, force)
    en = be.get_energies()
    d32 = s["displ"].astype(np.float32)
    for r in range(R):
        S = O.System(s["charge"], s["sigma"], s["epsilon"], s["box"] * scale[r], s["cutoff"], s["ewald_alpha"], s["excl"])
        p1 = posq[r, :n, :3].astype(np.float64)
        p2 = (posq[r, :n, :3] + d32).astype(np.float64)
        e1, _, f1 = S.nb_direct(p1)
        e2, _, f2 = S.nb_direct(p2)
        prm = sched[4 + 11 * r]
        sc = O.scalars(prm, e1, e2)
        f_ref = O.merge_ref(np.zeros_like(f1), f1, f2, sc["sp_ref"], prm[8])
        f_gpu = force_from_fixed(force.cpu().numpy()[r], n, P)
        assert abs(en[r, E_U1] - e1) <= 1e-6 * abs(e1), (r, en[r, E_U1], e1)
        assert abs(en[r, E_USC] - sc["u_sc"]) <= 5e-3, (r, en[r, E_USC], sc["u_sc"])
        assert rel_rms(f_gpu, f_ref) <= 1e-5, r
    be.close()


def test_three_groups_interleaved_indices_and_equal_displacements():
    """Ragged / irregular topology: three displaced groups of whole waters picked with non-contiguous atom indices
    (every 5th / 7th molecule of two separate slabs), two of them with the SAME displacement vector (pairs between
    them move together, i.e. belong to both states) and one with the opposite one; N is neither a multiple of 8 nor of
    32.  Also evaluates the partially overlapping exclusion structure of the synthetic chain ligand."""
    from atmmetaforce import synthetic
    from helpers import rel_rms
    s = synthetic.water_box(9000, n_lig=13, seed=17)
    n = s["pos"].shape[0]
    assert n % 8 != 0
    pos, box = s["pos"], s["box"]
    nl = int(round((n / 3) ** (1.0 / 3.0)))
    a = box[0] / nl
    lig = np.nonzero(np.abs(s["displ"]).sum(1) > 0)[0]
    first_water = lig.max() + 1
    mol = (np.arange(n) - first_water) // 3
    is_water = np.arange(n) >= first_water
    d_up = np.array([0.5 * a, 0.5 * a, (int(0.4 * nl) + 0.5) * a])
    g1 = is_water & (pos[:, 2] < 0.2 * box[2]) & (mol % 5 == 0)
    g2 = is_water & (pos[:, 2] > 0.25 * box[2]) & (pos[:, 2] < 0.4 * box[2]) & (mol % 7 == 0)
    # whole molecules only: take the decision of the oxygen (first atom of each triple)
    for g in (g1, g2):
        ox = g[first_water::3].copy()
        g[first_water:] = np.repeat(ox, 3)[: n - first_water]
    s["displ"][g1] = d_up           # same vector as ...
    s["displ"][g2] = d_up           # ... this group: one displacement class
    s["displ"][lig] = -s["displ"][lig]   # the ligand goes the other way
    assert g1.sum() > 20 and g2.sum() > 20
    params = [0.3, 0.6, 0.02, 50.0, 1.5, 1e7, 5e6, 0.0625, -1.0]     # alpha > 0, direction -1, w0 != 0
    res = _run(s, s["cutoff"], s["ewald_alpha"], params, skin=0.07)
    en = res["en"]
    print("three groups: U1 %.3f/%.3f u %.4f/%.4f frms %.2e stats %s" % (en[E_U1], res["e1"], en[E_USC], res["sc"]["u_sc"],
          rel_rms(res["f_gpu"], res["f_ref"]), res["stats"]))
    assert res["stats"]["groups"] == 2          # equal displacement vectors form ONE class
    assert abs(en[E_U1] - res["e1"]) <= 1e-6 * abs(res["e1"])
    assert abs(en[E_U2] - res["e2"]) <= 1e-6 * abs(res["e2"])
    # hundreds of displaced atoms: |u| ~ 1e5 kJ/mol, so the energy bar is the relative one (1e-6)
    assert abs(en[E_USC] - res["sc"]["u_sc"]) <= max(2e-2, 1e-6 * abs(res["sc"]["u_sc"]))
    assert abs(en[E_SP] - res["sc"]["sp"]) <= 1e-4
    assert rel_rms(res["f_gpu"], res["f_ref"]) <= 1e-5


def test_tiny_system_in_a_large_box():
    """37 atoms (a 13-atom ligand and 8 waters) in a 4.4 nm box: almost every cluster slot and list step is padding."""
    from atmmetaforce import synthetic
    from helpers import rel_rms
    s = synthetic.water_box(9000, n_lig=13, seed=17)
    keep = np.arange(37)
    idx = {int(a): k for k, a in enumerate(keep)}
    t = dict(s)
    for key in ("pos", "charge", "sigma", "epsilon", "displ"):
        t[key] = s[key][keep].copy()
    t["excl"] = np.array([[idx[a], idx[b]] for a, b in s["excl"] if a in idx and b in idx], np.int32).reshape(-1, 2)
    t["exc14"] = np.zeros((0, 2), np.int32)
    t["exc14_par"] = np.zeros((0, 3))
    # spread the waters so that some pairs are inside and some outside the cutoff
    t["pos"][13:] += np.random.default_rng(2).normal(0, 0.15, (24, 3))
    params = [0.5, 0.5, 0, 0, 0, 800, 400, 0.0625, 1.0]
    res = _run(t, s["cutoff"], s["ewald_alpha"], params, skin=0.1)
    en = res["en"]
    assert abs(en[E_U1] - res["e1"]) <= 1e-6 * abs(res["e1"]) + 1e-4
    assert abs(en[E_U2] - res["e2"]) <= 1e-6 * abs(res["e2"]) + 1e-4
    assert abs(en[E_USC] - res["sc"]["u_sc"]) <= 5e-3
    assert rel_rms(res["f_gpu"], res["f_ref"]) <= 1e-5


def test_config5_500k_water_box():
    """BASELINE configs[4], largest size: ~500k-atom water box with a 50-atom ligand, against the oracle (which still
    finishes in seconds at this size) plus two size-independent properties: Newton's third law on the merged force
    (the fixed-point sum over all atoms vanishes to rounding) and bit-identical results from a second evaluation."""
    from atmmetaforce import synthetic
    from helpers import rel_rms
    s = synthetic.water_box(500_000, n_lig=50)
    params = synthetic.atm_schedule_22()[16]          # direction -1 leg
    res = _run(s, s["cutoff"], s["ewald_alpha"], params, skin=0.1)
    err = _check(res, tol_u=5e-3)
    f = res["f_gpu"]
    assert np.abs(f.sum(0)).max() <= 1e-7 * np.abs(f).sum()
    print("500k: N %d u %.4f force rel rms %.2e |sum F|/sum|F| %.1e" % (f.shape[0], res["en"][E_U], err,
          np.abs(f.sum(0)).max() / np.abs(f).sum()))


def test_concurrent_prune_serves_the_next_step(abfe):
    """atm_step_io.concurrent_prune: the step that carries the prune still runs on the list in use (complete, so the
    result is right); the NEXT step runs on the list pruned from the carried coordinates -- bit-identical to
    atm_nb_prune + atm_step on those coordinates (same list, fixed-point accumulation)."""
    import torch
    import atmmetaforce as atm
    import oracle_py as O
    from helpers import make_backend, oracle_system, force_from_fixed, rel_rms
    alpha = O.ewald_alpha(1.0)
    rng = np.random.default_rng(8)
    n = abfe["pos"].shape[0]
    moved = abfe["pos"] + rng.normal(0, 0.01, abfe["pos"].shape)
    out = {}
    for mode in ("concurrent", "before"):
        be, posq0, _ = make_backend(atm, abfe, 1.0, alpha, abfe["params"], skin=0.1)
        P = be.P
        posq1 = posq0.clone()
        posq1[0, :n, :3] = torch.from_numpy(moved.astype(np.float32)).cuda()
        stream = torch.cuda.Stream()
        fa, fb = (torch.zeros((1, 3 * P), dtype=torch.int64, device="cuda") for _ in range(2))
        with torch.cuda.stream(stream):
            be.rebuild(posq0, stream=stream)
            if mode == "concurrent":
                be.step(posq1, fa, stream=stream, concurrent_prune=True, graph=True)   # old list, prune of posq1 alongside
                be.step(posq1, fb, stream=stream, graph=True)                          # new list
            else:
                be.step(posq1, fa, stream=stream)                                      # old list
                be.prune(posq1, stream=stream)
                be.step(posq1, fb, stream=stream)                                      # new list
        en = be.get_energies(stream=stream)[0].copy()
        out[mode] = (fa.cpu(), fb.cpu(), en)
        be.close()
    assert torch.equal(out["concurrent"][1], out["before"][1])
    assert torch.equal(out["concurrent"][0], out["before"][0])
    assert np.array_equal(out["concurrent"][2][:7], out["before"][2][:7])
    S = oracle_system(O, abfe, 1.0, alpha)
    x1 = moved.astype(np.float32).astype(np.float64)
    x2 = (moved.astype(np.float32) + abfe["displ"].astype(np.float32)).astype(np.float64)
    e1, _, f1 = S.nb_direct(x1)
    e2, _, f2 = S.nb_direct(x2)
    sc = O.scalars(abfe["params"], e1, e2)
    f_ref = O.merge_ref(np.zeros_like(f1), f1, f2, sc["sp_ref"], abfe["params"][8])
    for k in (0, 1):
        assert rel_rms(force_from_fixed(out["concurrent"][k].numpy()[0], n, be.P), f_ref) <= 1e-5


def test_async_rebuild_overflow_is_never_silent():
    """A pair list that outgrows its capacity during an ASYNCHRONOUS rebuild (a droplet contracting to several times its
    density after the verified first build sized the lists): the steps computed from the truncated lists return NaN
    energies and NaN-poisoned forces, atm_nb_check / the next API call report ATM_ERR_STATE, and the rebuild that
    follows reallocates, after which results match the oracle again.  A list that only comes close to its capacity
    makes the next rebuild reallocate ahead of time without any error."""
    import torch
    import atmmetaforce as atm
    import oracle_py as O
    from atmmetaforce import synthetic, _capi
    from helpers import oracle_system, rel_rms, force_from_fixed
    s = synthetic.water_box(20000, n_lig=12, seed=11)   # 5.8 nm box: the stretched clusters of the depleted shell still fit
    n = s["pos"].shape[0]
    params = synthetic.atm_schedule_22()[5]
    be = atm.ATMBackend(n, precision="mixed", num_replicas=1)
    P = be.P
    be.set_displacements(s["displ"])
    be.set_box(s["box"])
    be.set_parameters(params)
    be.nb_setup(s["charge"], s["sigma"], s["epsilon"], s["cutoff"], s["ewald_alpha"], skin=0.05, skin_outer=0.1, exclusions=s["excl"])
    stream = torch.cuda.Stream()

    def posq_of(pos):
        q = np.zeros((1, P, 4), np.float32)
        q[0, :n, :3] = pos
        q[0, :n, 3] = s["charge"]
        return torch.from_numpy(q).cuda()

    posq0 = posq_of(s["pos"])
    force = torch.zeros((1, 3 * P), dtype=torch.int64, device="cuda")
    with torch.cuda.stream(stream):
        be.rebuild(posq0, stream=stream)                      # first build: synchronous, verified, sizes the lists
        be.step(posq0, force, stream=stream)
    en0 = be.get_energies(stream=stream)[0].copy()
    assert np.isfinite(en0[_capi.E_U1])
    cap_before = be.nb_stats()["cap_env"]

    # contract everything inside a sphere towards its centre: ~4x the partners within the list radius of a central
    # cluster, far beyond the capacity margin of the verified first build (1.6x the mean-density expectation)
    c = 0.5 * np.asarray(s["box"], np.float64).reshape(-1)[:3]
    pos = s["pos"].copy()
    d = pos - c
    rr = np.linalg.norm(d, axis=1)
    inside = rr < 1.5
    pos[inside] = c + d[inside] * 0.55
    posq1 = posq_of(pos)
    force1 = torch.zeros_like(force)
    with torch.cuda.stream(stream):
        be.rebuild(posq1, stream=stream)                      # asynchronous now: truncates at least one list
        be.step(posq1, force1, stream=stream)
    stream.synchronize()
    with pytest.raises(atm.ATMError, match="capacity"):
        be.nb_check(wait=True)
    en_raw = torch.as_tensor(_DevView(be.energies_device_ptr(), (1, _capi.NUM_ENERGY_SLOTS), "<f8"), device="cuda").cpu().numpy()[0]
    assert np.isnan(en_raw[_capi.E_U1]) and np.isnan(en_raw[_capi.E_U]) and np.isnan(en_raw[_capi.E_ENERGY]) and np.isnan(en_raw[_capi.E_SP])
    f_bad = force_from_fixed(force1.cpu().numpy()[0], n, P)
    assert np.abs(f_bad).min() > 1e8                          # every atom's force is poisoned, none looks plausible
    with pytest.raises(atm.ATMError):                         # further steps are refused until the lists are rebuilt
        be.step(posq1, force1, stream=stream)

    # cure: rebuild again (reallocates with the raised capacity, synchronous + verified), repeat the step
    force2 = torch.zeros_like(force)
    with torch.cuda.stream(stream):
        be.rebuild(posq1, stream=stream)
        be.step(posq1, force2, stream=stream)
    en = be.get_energies(stream=stream)[0]
    be.nb_check(wait=True)
    assert be.nb_stats()["cap_env"] > cap_before
    S = oracle_system(O, s, s["cutoff"], s["ewald_alpha"])
    x1 = pos.astype(np.float32).astype(np.float64)
    x2 = (pos.astype(np.float32) + s["displ"].astype(np.float32)).astype(np.float64)
    e1, _, f1 = S.nb_direct(x1)
    e2, _, f2 = S.nb_direct(x2)
    sc = O.scalars(params, e1, e2)
    assert abs(en[_capi.E_U1] - e1) <= 1e-6 * abs(e1)
    f_ok = force_from_fixed(force2.cpu().numpy()[0], n, P)
    assert rel_rms(f_ok, O.merge_ref(np.zeros_like(f1), f1, f2, sc["sp_ref"], params[8])) <= 1e-5
    be.close()


class _DevView:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}


def test_own_sort_front_end_equals_cub_radix_sort(tmp_path):
    """The rebuild's own sort front end (bin -> scan -> in-warp bitonic sort + place -> pack + links + boxes) orders the
    sites exactly as the CUB radix sort it replaces (ties by site index), so clusters, lists, forces and energies are
    BIT-identical; ATM_B200_CUB_SORT=1 selects the CUB path (read once per process, hence two subprocesses)."""
    import os
    import subprocess
    import sys
    script = tmp_path / "run.py"
    script.write_text('''
import hashlib, os, sys
import numpy as np, torch
root = sys.argv[1]
sys.path.insert(0, os.path.join(root, "openmm-atmmetaforce-plugin_b200", "python"))
import atmmetaforce as atm
from atmmetaforce import synthetic
h = hashlib.sha256()
for name, s, R in (("config3", synthetic.config3(), 3), ("rbfe", dict(np.load(os.path.join(root, "tests", "golden", "temoa_g1_g4_rbfe.npz"))), 1)):
    if name == "rbfe":
        s["cutoff"], s["ewald_alpha"] = 1.0, synthetic.ewald_alpha(1.0)
    n = s["pos"].shape[0]
    be = atm.ATMBackend(n, precision="mixed", num_replicas=R)
    be.set_displacements(s["displ"]); be.set_box(s["box"])
    sched = synthetic.atm_schedule_22()
    for r in range(R):
        be.set_parameters(sched[(7 * r + 2) % 22], replica=r)
    be.nb_setup(s["charge"], s["sigma"], s["epsilon"], s["cutoff"], s["ewald_alpha"], skin=0.05, skin_outer=0.3, exclusions=s["excl"],
                exception_pairs=s["exc14"], exception_params=s["exc14_par"])
    rng = np.random.default_rng(4)
    posq = np.zeros((R, be.P, 4), np.float32)
    for r in range(R):
        posq[r, :n, :3] = s["pos"] + rng.normal(0, 0.004, (n, 3)) * (r > 0)
        posq[r, :n, 3] = s["charge"]
    posq = torch.from_numpy(posq).cuda()
    force = torch.zeros((R, 3 * be.P), dtype=torch.int64, device="cuda")
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        be.rebuild(posq, stream=st); be.step(posq, force, stream=st)
        posq[:, :n, :3] += 0.01
        be.rebuild(posq, stream=st); be.step(posq, force, stream=st, graph=True)     # asynchronous (graph) rebuild
    en = be.get_energies(stream=st)
    h.update(force.cpu().numpy().tobytes()); h.update(np.ascontiguousarray(en[:, :7]).tobytes())
    h.update(repr(sorted(be.nb_stats().items())).encode())
    be.close()
print("HASH", h.hexdigest())
''')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = {}
    for cub in ("0", "1"):
        r = subprocess.run([sys.executable, str(script), root], capture_output=True, text=True, timeout=600, env=dict(os.environ, ATM_B200_CUB_SORT=cub))
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        out[cub] = [ln for ln in r.stdout.splitlines() if ln.startswith("HASH")][0]
    assert out["0"] == out["1"]


def test_headline_workload_full_size_properties():
    """BASELINE configs[2] at its full size -- 22 replicas of the 23k-atom system, one lambda state each, in ONE handle, the
    bench's pair-list settings -- through properties that do not need the oracle at that size: every replica's merged force
    sums to zero over the atoms (Newton's third law in each state, fixed-point sums), a replica taken alone in a
    one-replica handle gives the same bits (the batching is invisible), the graph replay gives the same bits, and the
    replica with lambda = 0 feels exactly the state-1 force (sp = 0).  Three replicas are also checked against the oracle."""
    import torch
    import atmmetaforce as atm
    import oracle_py as O
    from atmmetaforce import synthetic
    from helpers import oracle_system, force_from_fixed, rel_rms
    s = synthetic.config3()
    n = s["pos"].shape[0]
    sched = synthetic.atm_schedule_22()
    R = 22

    def handle(rows):
        be = atm.ATMBackend(n, precision="mixed", num_replicas=len(rows))
        be.set_displacements(s["displ"])
        be.set_box(s["box"])
        for r, row in enumerate(rows):
            be.set_parameters(row, replica=r)
        be.nb_setup(s["charge"], s["sigma"], s["epsilon"], s["cutoff"], s["ewald_alpha"], skin=0.05, skin_outer=0.3,
                    exclusions=s["excl"], exception_pairs=s["exc14"], exception_params=s["exc14_par"])
        return be

    be = handle([sched[r] for r in range(R)])
    P = be.P
    rng = np.random.default_rng(77)
    posq = np.zeros((R, P, 4), np.float32)
    for r in range(R):
        posq[r, :n, :3] = s["pos"] + rng.normal(0, 0.005, (n, 3))
        posq[r, :n, 3] = s["charge"]
    d_posq = torch.from_numpy(posq).cuda()
    force = torch.zeros((R, 3 * P), dtype=torch.int64, device="cuda")
    be.rebuild(d_posq)
    be.step(d_posq, force)
    en = be.get_energies()
    f_all = force.cpu().numpy()
    assert np.isfinite(en[:, :7]).all()
    for r in range(R):
        f = force_from_fixed(f_all[r], n, P)
        assert np.abs(f.sum(0)).max() <= 1e-7 * np.abs(f).sum(), r
    # graph replay: same bits
    force_g = torch.zeros_like(force)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        be.step(d_posq, force_g, graph=True, stream=stream)
    stream.synchronize()
    assert torch.equal(force_g, force)
    # lambda = 0 end state (schedule row 0): sp = 0, the merged force is the state-1 force; lambda = 1 rows: sp = 1 or 0 by leg
    assert en[0, E_SP] == 0.0
    # one replica alone: same bits as inside the batch
    for r in (0, 7, 21):
        one = handle([sched[r]])
        f1 = torch.zeros((1, 3 * P), dtype=torch.int64, device="cuda")
        one.rebuild(d_posq[r:r + 1].contiguous())
        one.step(d_posq[r:r + 1].contiguous(), f1)
        assert torch.equal(f1[0], force[r]), r
        assert np.array_equal(one.get_energies()[0][:7], en[r][:7])
        one.close()
    # oracle on three of the replicas
    S = oracle_system(O, s, s["cutoff"], s["ewald_alpha"])
    d32 = s["displ"].astype(np.float32)
    for r in (3, 11, 18):
        p1 = posq[r, :n, :3].astype(np.float64)
        p2 = (posq[r, :n, :3] + d32).astype(np.float64)
        e1, _, g1 = S.nb_direct(p1)
        e2, _, g2 = S.nb_direct(p2)
        sc = O.scalars(sched[r], e1, e2)
        f_ref = O.merge_ref(np.zeros_like(g1), g1, g2, sc["sp_ref"], sched[r][8])
        assert abs(en[r, E_U1] - e1) <= 1e-6 * abs(e1)
        assert abs(en[r, E_USC] - sc["u_sc"]) <= 5e-3
        assert rel_rms(force_from_fixed(f_all[r], n, P), f_ref) <= 1e-5
    be.close()
