"""GPU parity of the two-state PME reciprocal space against the oracle's double-precision
smooth-PME restatement (same mesh, same spline order), and the reference pin with EVERYTHING evaluated on the GPU."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

E_U1, E_U2, E_U, E_USC, E_EBIAS, E_ENERGY, E_SP = range(7)
E_UREC1, E_UREC2, E_USELF = 11, 12, 13


def _run_pme(s, cutoff, alpha, params, grid, order=5):
    import torch
    import atmmetaforce as atm
    import oracle_py as O
    from helpers import oracle_system, make_backend, force_from_fixed
    n = s["pos"].shape[0]
    be, posq, _ = make_backend(atm, s, cutoff, alpha, params, skin=0.1)
    be.pme_setup(grid, order)
    be.rebuild(posq)
    force = torch.zeros((1, 3 * be.P), dtype=torch.int64, device="cuda")
    be.step(posq, force)
    en = be.get_energies()[0]
    f_gpu = force_from_fixed(force.cpu().numpy()[0], n, be.P)
    force2 = torch.zeros_like(force)
    be.step(posq, force2)
    torch.cuda.synchronize()
    assert torch.equal(force, force2)          # fixed-point spreading: bit-reproducible
    be.close()
    S = oracle_system(O, s, cutoff, alpha)
    pos1 = s["pos"].astype(np.float32).astype(np.float64)
    pos2 = (s["pos"].astype(np.float32) + s["displ"].astype(np.float32)).astype(np.float64)
    e1, _, f1 = S.nb_direct(pos1)
    e2, _, f2 = S.nb_direct(pos2)
    r1, g1 = S.pme_recip(pos1, grid, order, want_force=True)
    r2, g2 = S.pme_recip(pos2, grid, order, want_force=True)
    self_e = -138.935456 * alpha / np.sqrt(np.pi) * float((s["charge"] ** 2).sum())
    U1, U2 = e1 + r1 + self_e, e2 + r2 + self_e
    sc = O.scalars(params, U1, U2)
    f_ref = O.merge_ref(np.zeros_like(f1), f1 + g1, f2 + g2, sc["sp_ref"], params[8])
    return dict(en=en, f_gpu=f_gpu, f_ref=f_ref, U1=U1, U2=U2, r1=r1, r2=r2, self_e=self_e, sc=sc)


def test_abfe_pin_fully_on_gpu(abfe):
    """Direct space + PME reciprocal of both states on the GPU: u = 58.2 +- 0.1 (ref: python/tests/test_abfe.py:148-150)
    with OpenMM's mesh rule for tolerance 5e-4 (35 x 40 x 35 here)."""
    import oracle_py as O
    from helpers import rel_rms
    alpha = O.ewald_alpha(1.0)
    grid = O.pme_grid(abfe["box"], alpha)
    r = _run_pme(abfe, 1.0, alpha, abfe["params"], grid)
    en = r["en"]
    print("abfe+PME: u_sc %.4f  Urec1 %.4f (oracle %.4f)  dUrec %.5f (oracle %.5f)  U1 %.3f (oracle %.3f)  frms %.2e" % (
        en[E_USC], en[E_UREC1], r["r1"], en[E_UREC2] - en[E_UREC1], r["r2"] - r["r1"], en[E_U1], r["U1"], rel_rms(r["f_gpu"], r["f_ref"])))
    assert abs(en[E_USC] - 58.2) <= 0.1
    assert abs(en[E_UREC1] - r["r1"]) <= 1e-6 * abs(r["r1"]) + 1e-4
    assert abs((en[E_UREC2] - en[E_UREC1]) - (r["r2"] - r["r1"])) <= 1e-3    # float meshes (measured 4.7e-4): DESIGN.md section 3.4
    assert abs(en[E_USELF] - r["self_e"]) <= 1e-6 * abs(r["self_e"])
    assert abs(en[E_U1] - r["U1"]) <= 1e-6 * abs(r["U1"])
    assert abs(en[E_USC] - r["sc"]["u_sc"]) <= 5e-3
    assert rel_rms(r["f_gpu"], r["f_ref"]) <= 1e-5


def test_abfe_potential_energy_pin_on_gpu(abfe):
    """Reference pin no. 2 (python/tests/test_abfe.py:147,149): PE of force groups {0, ATM} = -116071.0 +- 0.1 kJ/mol.
    The ATM force's energy e0 + W (the complete NonbondedForce of state 1 incl. dispersion correction, plus lambda2 u)
    comes from the GPU; the group-0 terms (bonds, angles, torsions, restraints; not on the ATM path) from the numpy
    oracle.  Also checks that the dispersion switch changes U1 and U2 by the same constant and nothing else."""
    import torch
    import atmmetaforce as atm
    import oracle_py as O
    import oracle_bonded as B
    from helpers import make_backend
    alpha = O.ewald_alpha(1.0)
    grid = O.pme_grid(abfe["box"], alpha)
    be, posq, _ = make_backend(atm, abfe, 1.0, alpha, abfe["params"], skin=0.1)
    be.pme_setup(grid, 5)
    be.rebuild(posq)
    force0 = torch.zeros((1, 3 * be.P), dtype=torch.int64, device="cuda")
    be.step(posq, force0)
    en0 = be.get_energies()[0].copy()
    be.set_dispersion_correction(True)
    force = torch.zeros_like(force0)
    be.step(posq, force, graph=True, stream=torch.cuda.Stream())
    torch.cuda.synchronize()
    en = be.get_energies()[0].copy()
    be.close()
    disp = B.dispersion_correction(abfe["sigma"], abfe["epsilon"], 1.0, float(abfe["box"].prod()))
    assert abs((en[E_U1] - en0[E_U1]) - disp) <= 2e-3 and abs((en[E_U2] - en0[E_U2]) - disp) <= 2e-3
    assert en[E_USC] == en0[E_USC] and en[E_SP] == en0[E_SP] and torch.equal(force, force0)
    g0, _ = B.group0_energy(abfe)
    pe = g0 + en[E_ENERGY]
    print("abfe PE on the GPU: %.4f (pin %.1f), dispersion %.4f" % (pe, float(abfe["pin_pe"]), disp))
    assert abs(pe - float(abfe["pin_pe"])) <= 0.1


def test_rbfe_pme(rbfe):
    import oracle_py as O
    from helpers import rel_rms
    alpha = O.ewald_alpha(1.0)
    grid = O.pme_grid(rbfe["box"], alpha)
    r = _run_pme(rbfe, 1.0, alpha, rbfe["params"], grid)
    en = r["en"]
    assert abs(en[E_U1] - r["U1"]) <= 1e-6 * abs(r["U1"]) and abs(en[E_U2] - r["U2"]) <= 1e-6 * abs(r["U2"])
    assert abs(en[E_U] - (r["U2"] - r["U1"])) <= 5e-3
    assert abs(en[E_U] - 2.107) <= 0.05        # survey-time value with the exact Ewald sum (unpinned by the reference)
    assert rel_rms(r["f_gpu"], r["f_ref"]) <= 1e-5


def test_pme_order_and_odd_grid():
    """Spline order 4 and an odd, non-cubic mesh."""
    from atmmetaforce import synthetic
    from helpers import rel_rms
    s = synthetic.water_box(6000, n_lig=15, seed=6)
    params = synthetic.atm_schedule_22()[14]     # direction -1
    r = _run_pme(s, s["cutoff"], s["ewald_alpha"], params, [27, 30, 25], order=4)
    en = r["en"]
    # direct space, reciprocal space and the self energy nearly cancel in this system: the 1e-6 bar is relative to the
    # largest component (the self energy), not to their small sum
    assert abs(en[E_U1] - r["U1"]) <= 1e-6 * (abs(r["U1"]) + abs(r["self_e"]))
    assert abs(en[E_UREC1] - r["r1"]) <= 1e-6 * abs(r["r1"]) + 1e-4
    assert abs((en[E_UREC2] - en[E_UREC1]) - (r["r2"] - r["r1"])) <= 1e-3    # float meshes (measured 1.3e-4)
    assert rel_rms(r["f_gpu"], r["f_ref"]) <= 1e-5


def test_pme_order6_replicas_shifted_coordinates():
    """Spline order 6 on a mesh with a z extent divisible by 4 (the 128-bit row loads with three float4s), two replicas, the
    second with every coordinate shifted by whole box lengths and a fraction (unwrapped input, tiles that wrap around the
    periodic boundary): the replica results agree with the oracle and, up to the float rounding of the shifted coordinates,
    with each other."""
    import torch
    import atmmetaforce as atm
    import oracle_py as O
    from atmmetaforce import synthetic
    from helpers import oracle_system, make_backend, force_from_fixed, rel_rms
    s = synthetic.water_box(6000, n_lig=15, seed=9)
    params = synthetic.atm_schedule_22()[7]
    grid, order = [32, 36, 40], 6
    n = s["pos"].shape[0]
    be, posq, _ = make_backend(atm, s, s["cutoff"], s["ewald_alpha"], params, skin=0.1, replicas=2)
    be.set_parameters(params, replica=1)
    shift = np.array([2.0, -1.0, 3.0]) * s["box"] + np.array([0.37, 0.11, 0.05])
    posq[1, :n, :3] += torch.from_numpy(shift.astype(np.float32)).cuda()
    be.pme_setup(grid, order)
    be.rebuild(posq)
    force = torch.zeros((2, 3 * be.P), dtype=torch.int64, device="cuda")
    be.step(posq, force)
    en = be.get_energies()
    S = oracle_system(O, s, s["cutoff"], s["ewald_alpha"])
    x = posq.cpu().numpy()
    for r in range(2):
        pos1 = x[r, :n, :3].astype(np.float64)
        pos2 = (x[r, :n, :3] + s["displ"].astype(np.float32)).astype(np.float64)
        e1, _, f1 = S.nb_direct(pos1)
        e2, _, f2 = S.nb_direct(pos2)
        r1, g1 = S.pme_recip(pos1, grid, order, want_force=True)
        r2, g2 = S.pme_recip(pos2, grid, order, want_force=True)
        assert abs(en[r][E_UREC1] - r1) <= 1e-6 * abs(r1) + 1e-4
        assert abs((en[r][E_UREC2] - en[r][E_UREC1]) - (r2 - r1)) <= 1e-3
        sc = O.scalars(params, e1 + r1, e2 + r2)
        f_ref = O.merge_ref(np.zeros_like(f1), f1 + g1, f2 + g2, sc["sp_ref"], params[8])
        assert rel_rms(force_from_fixed(force.cpu().numpy()[r], n, be.P), f_ref) <= 1e-5
    assert abs(en[0][E_UREC1] - en[1][E_UREC1]) <= 1e-5 * abs(en[0][E_UREC1])
    be.close()


def test_pme_float_and_double_mesh_pipelines_agree(tmp_path):
    """ATM_B200_PME_F64=1 selects the double-precision mesh pipeline of round 1 (global fixed-point spread, D2Z / Z2D); the
    default single-precision pipeline must give the same energies and forces to float-mesh accuracy."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "pme_case.py"
    script.write_text('''
import json, sys, numpy as np, torch
root = sys.argv[1]
sys.path.insert(0, root + "/openmm-atmmetaforce-plugin_b200/python"); sys.path.insert(0, root + "/tests")
import atmmetaforce as atm
from atmmetaforce import synthetic
from helpers import make_backend
s = synthetic.water_box(9000, n_lig=20, seed=4)
be, posq, _ = make_backend(atm, s, s["cutoff"], s["ewald_alpha"], synthetic.atm_schedule_22()[9], skin=0.1)
be.pme_setup([40, 40, 40], 5)
be.rebuild(posq)
force = torch.zeros((1, 3 * be.P), dtype=torch.int64, device="cuda")
be.step(posq, force)
en = be.get_energies()[0]
f = force.cpu().numpy()[0].astype(np.float64) / 4294967296.0
np.save(sys.argv[2], f)
print(json.dumps([float(v) for v in en]))
''')
    res = {}
    for f64, cub in (("0", "0"), ("1", "0"), ("0", "1")):
        out = str(tmp_path / f"f_{f64}{cub}.npy")
        p = subprocess.run([sys.executable, str(script), root, out], capture_output=True, text=True, timeout=600,
                           env=dict(os.environ, ATM_B200_PME_F64=f64, ATM_B200_CUB_SORT=cub))
        assert p.returncode == 0, p.stdout + p.stderr
        res[f64 + cub] = (np.array(json.loads(p.stdout.strip().splitlines()[-1])), np.load(out))
    (e32, f32), (e64, f64_) = res["00"], res["10"]
    # the tile-owned spread walks the bins of the pair-list structure: the CUB radix-sort front end (the fallback of the
    # own sort) fills the same bins in the same order, so every bit of the result is the same
    assert np.array_equal(res["01"][0][:7], e32[:7]) and np.array_equal(res["01"][1], f32)
    assert abs(e32[E_UREC1] - e64[E_UREC1]) <= 1e-6 * abs(e64[E_UREC1]) + 1e-4
    assert abs((e32[E_UREC2] - e32[E_UREC1]) - (e64[E_UREC2] - e64[E_UREC1])) <= 1e-3
    assert abs(e32[E_USC] - e64[E_USC]) <= 1e-3
    assert np.sqrt(((f32 - f64_) ** 2).sum() / (f64_ ** 2).sum()) <= 5e-6


def test_pme_site_drift_beyond_the_skin_poisons_the_step():
    """The tile-owned spread finds sites through the xy column they were sorted into at the last rebuild.  A site that has
    moved further since then than the structure tolerates (half the outer skin beyond its cluster's bounding box) must
    not be silently dropped from the mesh: the step returns NaN energies until the next rebuild."""
    import torch
    import atmmetaforce as atm
    from atmmetaforce import synthetic
    from helpers import make_backend
    s = synthetic.water_box(9000, n_lig=20, seed=4)
    be, posq, _ = make_backend(atm, s, s["cutoff"], s["ewald_alpha"], synthetic.atm_schedule_22()[9], skin=0.1)
    be.pme_setup([40, 40, 40], 5)
    be.rebuild(posq)
    force = torch.zeros((1, 3 * be.P), dtype=torch.int64, device="cuda")
    be.step(posq, force)
    assert np.isfinite(be.get_energies()[0][:7]).all()
    moved = posq.clone()
    moved[0, 100, 0] += 0.04                       # within half the skin (0.05 nm here): still fine
    be.step(moved, force)
    assert np.isfinite(be.get_energies()[0][:7]).all()
    moved[0, 100, 0] += 1.0                        # far outside
    be.step(moved, force)
    assert np.isnan(be.get_energies()[0][E_U1]) and np.isnan(be.get_energies()[0][E_USC])
    with pytest.raises(atm.ATMError, match="moved further"):
        be.nb_check(wait=True)                     # the waiting check names the cause
    be.rebuild(moved)                              # the rebuild re-sorts the site: results again
    be.nb_check(wait=True)
    be.step(moved, force)
    assert np.isfinite(be.get_energies()[0][:7]).all()
    be.close()


def test_pme_headline_size_batch_is_invisible():
    """The bench's --pme workload at full size (22 replicas x 23k atoms, 60^3 mesh, one lambda state each): every replica of
    the batch agrees with the same replica evaluated alone (the transforms are batched differently, so agreement is to
    float-mesh accuracy, not to the bit), and a second step gives the same bits (fixed-point tiles, deterministic)."""
    import torch
    import atmmetaforce as atm
    from atmmetaforce import synthetic
    s = synthetic.config3()
    n = s["pos"].shape[0]
    sched = synthetic.atm_schedule_22()
    grid = synthetic.pme_grid(s["box"], s["ewald_alpha"])

    def handle(rows):
        be = atm.ATMBackend(n, precision="mixed", num_replicas=len(rows))
        be.set_displacements(s["displ"])
        be.set_box(s["box"])
        for r, row in enumerate(rows):
            be.set_parameters(row, replica=r)
        be.nb_setup(s["charge"], s["sigma"], s["epsilon"], s["cutoff"], s["ewald_alpha"], skin=0.05, skin_outer=0.3,
                    exclusions=s["excl"], exception_pairs=s["exc14"], exception_params=s["exc14_par"])
        be.pme_setup(grid)
        return be

    R = 22
    be = handle([sched[r] for r in range(R)])
    P = be.P
    rng = np.random.default_rng(5)
    posq = np.zeros((R, P, 4), np.float32)
    for r in range(R):
        posq[r, :n, :3] = s["pos"] + rng.normal(0, 0.005, (n, 3))
        posq[r, :n, 3] = s["charge"]
    d_posq = torch.from_numpy(posq).cuda()
    force = torch.zeros((R, 3 * P), dtype=torch.int64, device="cuda")
    be.rebuild(d_posq)
    be.step(d_posq, force)
    en = be.get_energies()
    assert np.isfinite(en[:, :14]).all() and (np.abs(en[:, E_UREC1]) > 1.0).all()
    force2 = torch.zeros_like(force)
    be.step(d_posq, force2)
    torch.cuda.synchronize()
    assert torch.equal(force, force2) and np.array_equal(be.get_energies()[:, :14], en[:, :14])
    for r in (2, 13):
        one = handle([sched[r]])
        f1 = torch.zeros((1, 3 * P), dtype=torch.int64, device="cuda")
        x = d_posq[r:r + 1].contiguous()
        one.rebuild(x)
        one.step(x, f1)
        e1 = one.get_energies()[0]
        assert abs(e1[E_UREC1] - en[r, E_UREC1]) <= 1e-6 * abs(en[r, E_UREC1])
        assert abs(e1[E_USC] - en[r, E_USC]) <= 1e-3
        a = f1[0].cpu().numpy().astype(np.float64)
        b = force[r].cpu().numpy().astype(np.float64)
        assert np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum()) <= 2e-6
        one.close()
    be.close()


def test_pme_per_replica_boxes():
    """Two replicas in ONE handle, the second in a 2 % larger box (barostat) with molecules translated by whole box vectors:
    the spread's column scan, the influence function, the blend and the gather all use the replica's own box."""
    import torch
    import atmmetaforce as atm
    import oracle_py as O
    from atmmetaforce import synthetic
    from helpers import force_from_fixed, rel_rms
    s = synthetic.water_box(9000, n_lig=30, seed=33)
    n = s["pos"].shape[0]
    sched = synthetic.atm_schedule_22()
    rng = np.random.default_rng(5)
    lig = np.nonzero(np.abs(s["displ"]).sum(1) > 0)[0]
    mol = np.full(n, -1)
    mol[lig] = 0
    rest = np.nonzero(mol < 0)[0]
    mol[rest] = 1 + np.arange(rest.size) // 3
    shifts = rng.integers(-1, 2, (mol.max() + 1, 3)).astype(np.float64)
    scale = [1.0, 1.02]
    grid, order = [40, 40, 40], 5
    be = atm.ATMBackend(n, precision="mixed", num_replicas=2)
    P = be.P
    be.set_displacements(s["displ"])
    for r in range(2):
        be.set_box(s["box"] * scale[r], replica=r)
        be.set_parameters(sched[4 + 11 * r], replica=r)
    be.nb_setup(s["charge"], s["sigma"], s["epsilon"], s["cutoff"], s["ewald_alpha"], skin=0.1, exclusions=s["excl"])
    be.pme_setup(grid, order)
    posq = np.zeros((2, P, 4), np.float32)
    for r in range(2):
        posq[r, :n, :3] = s["pos"] * scale[r] + shifts[mol] * (s["box"] * scale[r])
        posq[r, :n, 3] = s["charge"]
    d_posq = torch.from_numpy(posq).cuda()
    force = torch.zeros((2, 3 * P), dtype=torch.int64, device="cuda")
    be.rebuild(d_posq)
    be.step(d_posq, force)
    en = be.get_energies()
    d32 = s["displ"].astype(np.float32)
    for r in range(2):
        S = O.System(s["charge"], s["sigma"], s["epsilon"], s["box"] * scale[r], s["cutoff"], s["ewald_alpha"], s["excl"])
        p1 = posq[r, :n, :3].astype(np.float64)
        p2 = (posq[r, :n, :3] + d32).astype(np.float64)
        e1, _, f1 = S.nb_direct(p1)
        e2, _, f2 = S.nb_direct(p2)
        r1, g1 = S.pme_recip(p1, grid, order, want_force=True)
        r2, g2 = S.pme_recip(p2, grid, order, want_force=True)
        prm = sched[4 + 11 * r]
        sc = O.scalars(prm, e1 + r1, e2 + r2)
        f_ref = O.merge_ref(np.zeros_like(f1), f1 + g1, f2 + g2, sc["sp_ref"], prm[8])
        assert abs(en[r, E_UREC1] - r1) <= 1e-6 * abs(r1) + 1e-4, (r, en[r, E_UREC1], r1)
        assert abs((en[r, E_UREC2] - en[r, E_UREC1]) - (r2 - r1)) <= 1e-3
        assert rel_rms(force_from_fixed(force.cpu().numpy()[r], n, P), f_ref) <= 1e-5, r
    be.close()
