"""The reference's own test, restated against the drop-in Python API (ref: python/tests/test_abfe.py:22-150):
build the ATM force for the TEMOA-G1 system, set the nine parameters, evaluate, check the perturbation energy pin."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_BindingEnergy(abfe):
    import atmmetaforce as atm
    import oracle_py as O
    from helpers import oracle_system
    kcal = 4.184
    lmbd = 0.5
    lambda1, lambda2, alpha, u0, w0coeff = lmbd, lmbd, 0.0, 0.0, 0.0
    umsc, ubcore, acore, direction = 200.0 * kcal, 100.0 * kcal, 0.0625, 1.0
    displ = [2.2, 2.2, 2.2]   # 22 Angstrom
    n = abfe["pos"].shape[0]
    lig_atoms = abfe["lig1"]

    atmforcegroup, nonbonded_force_group = 2, 1
    nonbonded = atm.NonbondedDirect(abfe["charge"], abfe["sigma"], abfe["epsilon"], cutoff=1.0, ewald_tolerance=5e-4,
                                    exclusions=abfe["excl"], exception_pairs=abfe["exc14"], exception_params=abfe["exc14_par"],
                                    force_group=nonbonded_force_group, reciprocal_space=False, dispersion_correction=False)
    atmforce = atm.ATMMetaForce(lambda1, lambda2, alpha, u0, w0coeff, umsc, ubcore, acore, direction, [nonbonded_force_group])
    for i in range(n):
        atmforce.addParticle(i, 0., 0., 0.)
    for i in lig_atoms:
        atmforce.setParticleParameters(int(i), int(i), displ[0], displ[1], displ[2])
    atmforce.setForceGroup(atmforcegroup)

    context = atm.Context(atmforce, nonbonded, abfe["box"], precision="mixed")
    context.setPositions(abfe["pos"])
    # override ATM parameters as the reference test does after loadState (:131-139; it sets ATMDirection to acore > 0)
    for name, val in ((atmforce.Lambda1(), lambda1), (atmforce.Lambda2(), lambda2), (atmforce.Alpha(), alpha),
                      (atmforce.U0(), u0), (atmforce.W0(), w0coeff), (atmforce.Umax(), umsc), (atmforce.Ubcore(), ubcore),
                      (atmforce.Acore(), acore), (atmforce.Direction(), acore)):
        context.setParameter(name, val)
    # PME reciprocal space stays in OpenMM's inner contexts; here the oracle's exact Ewald sum plays that role
    S = oracle_system(O, abfe, 1.0, nonbonded.ewald_alpha)
    r1, _ = S.ewald_recip(abfe["pos"], 1e-10)
    r2, _ = S.ewald_recip(abfe["pos"] + abfe["displ"], 1e-10)
    context.setExternalStateEnergies(r1, r2)

    state = context.getState(getEnergy=True, getForces=True, groups={0, atmforcegroup})
    pert_energy = atmforce.getPerturbationEnergy(context)
    assert np.allclose(58.2, pert_energy, atol=0.1), pert_energy
    # energy returned by the force = e0 + W(u) with lambda1 = lambda2 = 1/2, alpha = 0:  U1 + u/2
    e1, _, _ = S.nb_direct(abfe["pos"].astype(np.float32).astype(np.float64), want_force=False)
    assert abs(state.getPotentialEnergy() - (e1 + r1 + 0.5 * pert_energy)) < 1e-6 * abs(e1)
    assert state.getForces().shape == (n, 3)
    # the ATM group is skipped when it is not in `groups` (ref: ATMMetaForceImpl.cpp:103)
    assert context.getState(getEnergy=True, groups={0}).getPotentialEnergy() == 0.0

    # updateParametersInContext: zero displacement -> u == 0
    for i in lig_atoms:
        atmforce.setParticleParameters(int(i), int(i), 0., 0., 0.)
    atmforce.updateParametersInContext(context)
    context.setExternalStateEnergies(0.0, 0.0)
    context.getState(getEnergy=True)
    assert atmforce.getPerturbationEnergy(context) == 0.0
    context.close()


def test_whole_nonbonded_force_by_default(abfe):
    """Reference semantics by default (ref: copysystem keeps the reciprocal space in the cloned NonbondedForce,
    openmmapi/src/ATMMetaForceImpl.cpp:60-62; OpenMM's dispersion correction is on by default): the Python Context and the
    C++ Impl evaluate direct + reciprocal space + dispersion correction of both states without any external input, and
    reproduce both reference pins (python/tests/test_abfe.py:147-150)."""
    import atmmetaforce as atm
    from atmmetaforce import _atmmetaforce_core as core
    import oracle_bonded as B
    kcal = 4.184
    params = (0.5, 0.5, 0.0, 0.0, 0.0, 200.0 * kcal, 100.0 * kcal, 0.0625, 1.0)
    n = abfe["pos"].shape[0]
    g0, _ = B.group0_energy(abfe)

    def make_force():
        f = atm.ATMMetaForce(*params, [1])
        for i in range(n):
            f.addParticle(i, *abfe["displ"][i])
        f.setForceGroup(2)
        return f

    nonbonded = atm.NonbondedDirect(abfe["charge"], abfe["sigma"], abfe["epsilon"], cutoff=1.0, ewald_tolerance=5e-4,
                                    exclusions=abfe["excl"], exception_pairs=abfe["exc14"], exception_params=abfe["exc14_par"],
                                    force_group=1)
    fp = make_force()
    pctx = atm.Context(fp, nonbonded, abfe["box"], precision="mixed")
    pctx.setPositions(abfe["pos"])
    st = pctx.getState(getEnergy=True, getForces=True)
    assert abs(fp.getPerturbationEnergy(pctx) - 58.2) <= 0.1
    assert abs(g0 + st.getPotentialEnergy() - float(abfe["pin_pe"])) <= 0.1

    s = core.System()
    for m in abfe["mass"]:
        s.addParticle(float(m))
    L = abfe["box"]
    s.setDefaultPeriodicBoxVectors([L[0], 0, 0], [0, L[1], 0], [0, 0, L[2]])
    exc = {(int(a), int(b)): (0.0, 0.3, 0.0) for a, b in abfe["excl"]}
    for (a, b), p in zip(abfe["exc14"], abfe["exc14_par"]):
        exc[(int(a), int(b))] = tuple(float(x) for x in p)
    s.addNonbondedForce(abfe["charge"].tolist(), abfe["sigma"].tolist(), abfe["epsilon"].tolist(), [x for ab in exc for x in ab],
                        [x for ab in exc for x in exc[ab]], cutoff=1.0, ewaldTolerance=5e-4, forceGroup=1)
    fc = s.addATMMetaForce(make_force())
    ctx = core.Context(s)
    ctx.setPositions(abfe["pos"])
    e_cpp, f_cpp = ctx.calcForcesAndEnergy(True, True, 1 << 2)
    assert abs(core.ATMMetaForce.getPerturbationEnergy(fc, ctx) - 58.2) <= 0.1
    assert abs(g0 + e_cpp - float(abfe["pin_pe"])) <= 0.1
    # the two contexts use different pair-list skins (0.05 / 0.1 nm): the same pairs are summed in a different order
    from helpers import rel_rms
    assert abs(e_cpp - st.getPotentialEnergy()) <= 1e-8 * abs(e_cpp)
    assert rel_rms(f_cpp, st.getForces()) <= 2e-6
    pctx.close()


def test_kernel_object_matches_reference_seam(abfe):
    """ATMMetaForceB200Kernel (C++ facade object with the five methods of CalcATMMetaForceKernel) via pybind11."""
    import torch
    import atmmetaforce as atm
    from atmmetaforce import _atmmetaforce_core as core
    import oracle_py as O
    n = 1000
    P = 1024
    f = atm.ATMMetaForce(0.2, 0.6, 0.05, 50.0, 1.0, 800.0, 400.0, 0.0625, 1.0, [1])
    rng = np.random.default_rng(0)
    d = np.zeros((n, 3)); d[100:121] = [2.2, 2.2, 2.2]
    for i in range(n):
        f.addParticle(i, *d[i])
    perm = rng.permutation(n).astype(np.int32)
    k = core.ATMMetaForceB200Kernel()
    assert k.Name() == "CalcATMMetaForce"
    k.initialize(f, P, 1, perm.tolist(), 0)
    posq = torch.from_numpy(rng.uniform(-3, 6, (P, 4)).astype(np.float32)).cuda()
    corr = torch.zeros_like(posq)
    p1, p2, c1, c2 = (torch.zeros_like(posq) for _ in range(4))
    k.copyState(posq.data_ptr(), corr.data_ptr(), p1.data_ptr(), c1.data_ptr(), p2.data_ptr(), c2.data_ptr(), 0)
    torch.cuda.synchronize()
    table = O.displ_table(n, P, perm, d)
    e1, _, e2, _ = O.copy_state_f32(posq.cpu().numpy()[:n], corr.cpu().numpy()[:n], table[:n])
    assert np.array_equal(p2.cpu().numpy()[:n].view(np.uint32), e2.view(np.uint32))
    params = core.ATMMetaForceB200Kernel.getDefaultParameters(f)
    F = lambda: torch.from_numpy((rng.normal(0, 1000, 3 * P) * 2.0 ** 32).astype(np.int64)).cuda()
    f0, f1, f2 = F(), F(), F()
    f0c = f0.clone()
    energy = k.execute(params, f0.data_ptr(), f1.data_ptr(), f2.data_ptr(), -1000.0, -900.0, True, True, 0)
    torch.cuda.synchronize()
    sc = O.scalars([params[x] for x in atm.Context.PARAM_ORDER], -1000.0, -900.0)
    assert abs(energy - sc["energy"]) < 1e-9 and abs(k.getPerturbationEnergy() - sc["u_sc"]) < 1e-9
    exp = O.hybrid_force_i64(n, P, f0c.cpu().numpy(), f1.cpu().numpy(), f2.cpu().numpy(), sc["sp"])
    assert np.array_equal(f0.cpu().numpy(), exp)


def test_cpp_impl_matches_python_context_and_oracle(abfe):
    """The C++ orchestration (ATMMetaForceImpl behind the OpenMM-free System / Context, evaluating through
    atm_host_pipeline_step) on the reference fixture: same numbers as the Python Context on device buffers (both drive
    the same kernels with the same pair-list settings) and the oracle's direct-space energies and forces."""
    import atmmetaforce as atm
    from atmmetaforce import _atmmetaforce_core as core, _capi
    import oracle_py as O
    from helpers import oracle_system, rel_rms
    kcal = 4.184
    n = abfe["pos"].shape[0]
    params = (0.5, 0.5, 0.0, 0.0, 0.0, 200.0 * kcal, 100.0 * kcal, 0.0625, 1.0)
    atm_group, nb_group = 2, 1

    def make_force():
        f = atm.ATMMetaForce(*params, [nb_group])
        for i in range(n):
            f.addParticle(i, *abfe["displ"][i])
        f.setForceGroup(atm_group)
        return f

    # C++ path
    s = core.System()
    for m in abfe["mass"]:
        s.addParticle(float(m))
    L = abfe["box"]
    s.setDefaultPeriodicBoxVectors([L[0], 0, 0], [0, L[1], 0], [0, 0, L[2]])
    # OpenMM convention: every exception excludes its pair; non-zero parameters make it a 1-4 interaction
    exc = {(int(a), int(b)): (0.0, 0.3, 0.0) for a, b in abfe["excl"]}
    for (a, b), p in zip(abfe["exc14"], abfe["exc14_par"]):
        exc[(int(a), int(b))] = tuple(float(x) for x in p)
    pairs = [x for ab in exc for x in ab]
    pars = [x for ab in exc for x in exc[ab]]
    s.addNonbondedForce(abfe["charge"].tolist(), abfe["sigma"].tolist(), abfe["epsilon"].tolist(), pairs, pars,
                        cutoff=1.0, ewaldTolerance=5e-4, forceGroup=nb_group,
                        reciprocalSpaceForceGroup=31, useDispersionCorrection=False)   # direct space only: group 31 is not variable
    fc = s.addATMMetaForce(make_force())
    ctx = core.Context(s)
    ctx.setPairListSkins(fc, 0.1, 0.3)
    ctx.setPositions(abfe["pos"])
    e_cpp, f_cpp = ctx.calcForcesAndEnergy(True, True, -1)
    u_cpp = core.ATMMetaForce.getPerturbationEnergy(fc, ctx)
    rec = np.array(ctx.getEnergyRecord(fc))

    # Python Context on device buffers, same skins
    nonbonded = atm.NonbondedDirect(abfe["charge"], abfe["sigma"], abfe["epsilon"], cutoff=1.0, ewald_tolerance=5e-4,
                                    exclusions=abfe["excl"], exception_pairs=abfe["exc14"], exception_params=abfe["exc14_par"],
                                    force_group=nb_group, reciprocal_space=False, dispersion_correction=False)
    fp = make_force()
    pctx = atm.Context(fp, nonbonded, abfe["box"], precision="mixed", skin=0.1, skin_outer=0.3)
    pctx.setPositions(abfe["pos"])
    st = pctx.getState(getEnergy=True, getForces=True)
    # same kernels, same pair-list settings, fixed-point accumulation: equal up to the order of the excluded / 1-4 pair
    # lists (their double-precision energy partial sums are grouped per warp), i.e. far below any physical tolerance
    assert abs(e_cpp - st.getPotentialEnergy()) <= 1e-11 * abs(e_cpp)
    assert abs(u_cpp - fp.getPerturbationEnergy(pctx)) <= 1e-8
    assert np.abs(f_cpp - st.getForces()).max() <= 1e-6

    # oracle (direct space of both states on the float-rounded coordinates)
    S = oracle_system(O, abfe, 1.0, nonbonded.ewald_alpha)
    pos32 = abfe["pos"].astype(np.float32).astype(np.float64)
    pos2_32 = (abfe["pos"].astype(np.float32) + abfe["displ"].astype(np.float32)).astype(np.float64)
    e1, _, f1 = S.nb_direct(pos32)
    e2, _, f2 = S.nb_direct(pos2_32)
    sc = O.scalars(params, e1, e2)
    assert abs(rec[_capi.E_U1] - e1) <= 1e-6 * abs(e1)
    assert abs(u_cpp - sc["u_sc"]) <= 5e-3
    assert abs(e_cpp - sc["energy"]) <= 1e-6 * abs(sc["energy"])
    assert rel_rms(f_cpp, O.merge_ref(np.zeros_like(f1), f1, f2, sc["sp_ref"], params[8])) <= 1e-5

    # a second evaluation of moved coordinates (prune path), a parameter change, and a group mask without the ATM group
    rng = np.random.default_rng(5)
    moved = abfe["pos"] + rng.normal(0, 0.012, abfe["pos"].shape)   # > skin/2 for some atom, < skin_outer/2: prune
    ctx.setPositions(moved)
    ctx.setParameter("ATMLambda2", 0.75)
    e2_cpp, f2_cpp = ctx.calcForcesAndEnergy(True, True, 1 << atm_group)
    pctx.setPositions(moved)
    pctx.setParameter("ATMLambda2", 0.75)
    st2 = pctx.getState(getEnergy=True, getForces=True)
    assert abs(e2_cpp - st2.getPotentialEnergy()) <= 1e-9 * abs(e2_cpp)
    assert rel_rms(f2_cpp, st2.getForces()) <= 1e-6
    assert ctx.calcForcesAndEnergy(True, True, 1 << nb_group)[0] == 0.0

    # updateParametersInContext: zero displacements -> u == 0 (ref test idea: python/tests/test_abfe.py + ATMMetaForce.cpp:38-40)
    for i in abfe["lig1"]:
        fc.setParticleParameters(int(i), int(i), 0.0, 0.0, 0.0)
    core.ATMMetaForce.updateParametersInContext(fc, ctx)
    ctx.calcForcesAndEnergy(True, True, -1)
    assert core.ATMMetaForce.getPerturbationEnergy(fc, ctx) == 0.0
    pctx.close()


def test_cpp_impl_test_binary_gpu():
    """openmmapi/tests/TestATMMetaForceImpl.cpp in its gpu mode: a pure C++ host (no Python, no CUDA headers) builds a
    System, evaluates it through ATMMetaForceImpl, and checks translation invariance and u == 0 at zero displacement."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "openmm-atmmetaforce-plugin_b200", "build", "TestATMMetaForceImpl")
    if not os.path.exists(exe):
        pytest.skip("C++ test binary not built (run __graft_entry__.build())")
    out = subprocess.run([exe, "gpu"], capture_output=True, text=True)
    assert out.returncode == 0 and "Done" in out.stdout, out.stdout + out.stderr
