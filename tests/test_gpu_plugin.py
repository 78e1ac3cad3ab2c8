"""The reference's own orchestration on the GPU, through the compiled plugin glue: ATMMetaForceImpl on the "CUDA"
platform clones the NonbondedForce into two linked inner contexts, B200CalcATMMetaForceKernel (created by
Platform::createKernel("CalcATMMetaForce") from the factory libATMMetaForcePluginCUDA.so registered) does copyState and
the hybrid merge on the contexts' own device buffers (ref: openmmapi/src/ATMMetaForceImpl.cpp:90-128,
platforms/common/src/CommonATMMetaForceKernels.cpp:111-226)."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "openmm-atmmetaforce-plugin_b200")
PLUGIN = os.path.join(PKG, "libATMMetaForcePluginCUDA.so")


def test_plugin_cpp_binary_gpu():
    exe = os.path.join(PKG, "build", "TestB200ATMMetaForcePlugin")
    out = subprocess.run([exe, "gpu", PKG], capture_output=True, text=True)
    assert out.returncode == 0 and "Done" in out.stdout, out.stdout + out.stderr


def _cuda_context(core, abfe, params, nb_group, atm_group, precision="mixed"):
    import atmmetaforce as atm
    if "CUDA" not in core.getPlatformNames():
        core.registerCudaPlatform()
    core.loadPluginLibrary(PLUGIN)
    n = abfe["pos"].shape[0]
    f = atm.ATMMetaForce(*params, [nb_group])
    for i in range(n):
        f.addParticle(i, *abfe["displ"][i])
    f.setForceGroup(atm_group)
    s = core.System()
    for m in abfe["mass"]:
        s.addParticle(float(m))
    L = abfe["box"]
    s.setDefaultPeriodicBoxVectors([L[0], 0, 0], [0, L[1], 0], [0, 0, L[2]])
    exc = {(int(a), int(b)): (0.0, 0.3, 0.0) for a, b in abfe["excl"]}
    for (a, b), p in zip(abfe["exc14"], abfe["exc14_par"]):
        exc[(int(a), int(b))] = tuple(float(x) for x in p)
    s.addNonbondedForce(abfe["charge"].tolist(), abfe["sigma"].tolist(), abfe["epsilon"].tolist(), [x for ab in exc for x in ab],
                        [x for ab in exc for x in exc[ab]], cutoff=1.0, ewaldTolerance=5e-4, forceGroup=nb_group)
    fc = s.addATMMetaForce(f)
    ctx = core.Context(s, "CUDA", {"Precision": precision})
    return s, fc, ctx


def test_reference_orchestration_reproduces_the_pins(abfe):
    """Both golden vectors of the reference (python/tests/test_abfe.py:147-150) through the kernel seam: the inner
    contexts evaluate the complete NonbondedForce (direct + reciprocal space + dispersion correction) of each state."""
    from atmmetaforce import _atmmetaforce_core as core, _capi
    import oracle_py as O
    import oracle_bonded as B
    from helpers import oracle_system, rel_rms
    kcal = 4.184
    params = (0.5, 0.5, 0.0, 0.0, 0.0, 200.0 * kcal, 100.0 * kcal, 0.0625, 1.0)
    atm_group, nb_group = 2, 1
    s, fc, ctx = _cuda_context(core, abfe, params, nb_group, atm_group)
    assert ctx.getPlatformName() == "CUDA" and ctx.usesPlatformKernel(fc)
    n = abfe["pos"].shape[0]
    ctx.setPositions(abfe["pos"])
    e, f = ctx.calcForcesAndEnergy(True, True, 1 << atm_group)
    u = core.ATMMetaForce.getPerturbationEnergy(fc, ctx)
    rec = np.array(ctx.getEnergyRecord(fc))
    assert abs(u - 58.2) <= 0.1                                              # pin 1
    g0, _ = B.group0_energy(abfe)
    assert abs(g0 + e - float(abfe["pin_pe"])) <= 0.1                        # pin 2: PE of groups {0, ATM}

    # copyState wrote the inner coordinates bit-exactly (float add of the float-rounded displacement)
    p0, p1, p2 = (ctx.getInnerPosq(fc, k) for k in (0, 1, 2))
    table = O.displ_table(n, p0.shape[0], None, abfe["displ"])
    e1, _, e2, _ = O.copy_state_f32(p0[:n], np.zeros_like(p0[:n]), table[:n])
    assert np.array_equal(p1[:n].view(np.uint32), e1.view(np.uint32))
    assert np.array_equal(p2[:n].view(np.uint32), e2.view(np.uint32))

    # oracle: full NonbondedForce of both states at the float-rounded coordinates, merged with the Reference-platform rule
    alpha = O.ewald_alpha(1.0)
    grid = O.pme_grid(abfe["box"], alpha)
    S = oracle_system(O, abfe, 1.0, alpha)
    x1, x2 = e1[:, :3].astype(np.float64), e2[:, :3].astype(np.float64)
    d1, _, f1 = S.nb_direct(x1)
    d2, _, f2 = S.nb_direct(x2)
    r1, g1 = S.pme_recip(x1, grid, 5, want_force=True)
    r2, g2 = S.pme_recip(x2, grid, 5, want_force=True)
    const = -138.935456 * alpha / np.sqrt(np.pi) * float((abfe["charge"] ** 2).sum()) + \
        B.dispersion_correction(abfe["sigma"], abfe["epsilon"], 1.0, float(abfe["box"].prod()))
    U1, U2 = d1 + r1 + const, d2 + r2 + const
    sc = O.scalars(params, U1, U2)
    assert abs(rec[_capi.E_U1] - U1) <= 1e-6 * abs(U1) and abs(rec[_capi.E_U2] - U2) <= 1e-6 * abs(U2)
    assert abs(e - sc["energy"]) <= 1e-6 * abs(sc["energy"])
    f_ref = O.merge_ref(np.zeros_like(f1), f1 + g1, f2 + g2, sc["sp_ref"], params[8])
    assert rel_rms(f, f_ref) <= 1e-5

    # the variable group evaluated directly in the OUTER context is the state-1 NonbondedForce (same kernel, same numbers)
    e_nb, f_nb = ctx.calcForcesAndEnergy(True, True, 1 << nb_group)
    assert abs(e_nb - rec[_capi.E_U1]) <= 1e-9 * abs(e_nb)
    assert ctx.calcForcesAndEnergy(True, True, 1 << 5)[0] == 0.0

    # OpenMM re-sorts atoms: the reorder listeners re-upload the displacement table (and the stand-in NonbondedForce its
    # own order); every observable is unchanged and copyState stays bit-exact in the new slot order
    perm = np.random.default_rng(3).permutation(n).astype(np.int32)
    ctx.reorderAtoms(perm.tolist())
    assert ctx.getAtomIndex() == perm.tolist()
    e_b, f_b = ctx.calcForcesAndEnergy(True, True, 1 << atm_group)
    assert abs(e_b - e) <= 1e-9 * abs(e) and np.abs(f_b - f).max() <= 1e-6
    p0, p2 = ctx.getInnerPosq(fc, 0), ctx.getInnerPosq(fc, 2)
    table = O.displ_table(n, p0.shape[0], perm, abfe["displ"])
    _, _, e2b, _ = O.copy_state_f32(p0[:n], np.zeros_like(p0[:n]), table[:n])
    assert np.array_equal(p2[:n].view(np.uint32), e2b.view(np.uint32))

    # parameters are read from the context at every evaluation; direction -1 swaps the roles of the states
    ctx.setParameter("ATMDirection", -1.0)
    ctx.setParameter("ATMLambda1", 0.1)
    ctx.setParameter("ATMLambda2", 0.4)
    e_c, f_c = ctx.calcForcesAndEnergy(True, True, 1 << atm_group)
    p_c = (0.1, 0.4) + params[2:8] + (-1.0,)
    sc_c = O.scalars(p_c, U1, U2)
    assert abs(e_c - sc_c["energy"]) <= 1e-6 * abs(sc_c["energy"])
    assert abs(core.ATMMetaForce.getPerturbationEnergy(fc, ctx) - sc_c["u_sc"]) <= 5e-3
    assert rel_rms(f_c, O.merge_ref(np.zeros_like(f1), f1 + g1, f2 + g2, sc_c["sp_ref"], -1.0)) <= 1e-5


def test_fused_host_path_and_kernel_path_agree(abfe):
    """Same System on the kernel-less host platform (ONE fused two-state launch + two-state PME) and on the CUDA platform
    (two inner evaluations + kernel seam): same perturbation energy, same energy, same forces.  With the reciprocal
    space moved to a non-variable force group the fused path leaves it out, and u changes by exactly that difference."""
    from atmmetaforce import _atmmetaforce_core as core
    import oracle_py as O
    from helpers import oracle_system, rel_rms
    kcal = 4.184
    params = (0.5, 0.5, 0.0, 0.0, 0.0, 200.0 * kcal, 100.0 * kcal, 0.0625, 1.0)
    s, fc, ctx = _cuda_context(core, abfe, params, 1, 2)
    ctx.setPositions(abfe["pos"])
    e_k, f_k = ctx.calcForcesAndEnergy(True, True, 1 << 2)
    u_kernel = core.ATMMetaForce.getPerturbationEnergy(fc, ctx)
    host = core.Context(s)
    assert host.getPlatformName() == "HostB200" and not host.usesPlatformKernel(fc)
    host.setPositions(abfe["pos"])
    e_h, f_h = host.calcForcesAndEnergy(True, True, 1 << 2)
    u_fused = core.ATMMetaForce.getPerturbationEnergy(fc, host)
    assert abs(u_kernel - u_fused) <= 5e-3
    assert abs(e_k - e_h) <= 1e-6 * abs(e_h)
    assert rel_rms(f_k, f_h) <= 1e-5

    import atmmetaforce as atm
    n = abfe["pos"].shape[0]
    s2 = core.System()
    for m in abfe["mass"]:
        s2.addParticle(float(m))
    L = abfe["box"]
    s2.setDefaultPeriodicBoxVectors([L[0], 0, 0], [0, L[1], 0], [0, 0, L[2]])
    exc = {(int(a), int(b)): (0.0, 0.3, 0.0) for a, b in abfe["excl"]}
    for (a, b), p in zip(abfe["exc14"], abfe["exc14_par"]):
        exc[(int(a), int(b))] = tuple(float(x) for x in p)
    s2.addNonbondedForce(abfe["charge"].tolist(), abfe["sigma"].tolist(), abfe["epsilon"].tolist(), [x for ab in exc for x in ab],
                         [x for ab in exc for x in exc[ab]], cutoff=1.0, ewaldTolerance=5e-4, forceGroup=1, reciprocalSpaceForceGroup=31)
    f2 = atm.ATMMetaForce(*params, [1])
    for i in range(n):
        f2.addParticle(i, *abfe["displ"][i])
    f2.setForceGroup(2)
    fc2 = s2.addATMMetaForce(f2)
    host2 = core.Context(s2)
    host2.setPositions(abfe["pos"])
    host2.calcForcesAndEnergy(True, True, 1 << 2)
    u_direct = core.ATMMetaForce.getPerturbationEnergy(fc2, host2)
    alpha = O.ewald_alpha(1.0)
    S = oracle_system(O, abfe, 1.0, alpha)
    grid = O.pme_grid(abfe["box"], alpha)
    x1 = abfe["pos"].astype(np.float32).astype(np.float64)
    x2 = (abfe["pos"].astype(np.float32) + abfe["displ"].astype(np.float32)).astype(np.float64)
    r1, _ = S.pme_recip(x1, grid, 5)
    r2, _ = S.pme_recip(x2, grid, 5)
    assert abs((u_fused - u_direct) - (r2 - r1)) <= 5e-3


def _bonded_system(core, abfe, params, k_bond, var_groups):
    """NonbondedForce (group 1) + a HarmonicBondForce tying ligand atoms to host atoms (group 3) + ATMMetaForce (group 2)."""
    import atmmetaforce as atm
    n = abfe["pos"].shape[0]
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
    lig = [int(i) for i in abfe["lig1"][:3]]
    host = [5, 60, 130]
    bonds = [(a, b, 0.8 * float(np.linalg.norm(abfe["pos"][b] - abfe["pos"][a])), k_bond) for a, b in zip(lig, host)]
    if k_bond is not None:
        s.addHarmonicBondForce([b[0] for b in bonds], [b[1] for b in bonds], [b[2] for b in bonds], [b[3] for b in bonds], forceGroup=3)
    f = atm.ATMMetaForce(*params, list(var_groups))
    for i in range(n):
        f.addParticle(i, *abfe["displ"][i])
    f.setForceGroup(2)
    return s, s.addATMMetaForce(f), bonds


def _bonds_numpy(pos, bonds):
    e, f = 0.0, np.zeros_like(pos)
    for a, b, r0, k in bonds:
        d = pos[b] - pos[a]
        r = np.linalg.norm(d)
        e += 0.5 * k * (r - r0) ** 2
        f[a] += k * (r - r0) * d / r
        f[b] -= k * (r - r0) * d / r
    return e, f


def test_generic_variable_force_through_both_impl_modes(abfe):
    """A second Force in the variable force groups (HarmonicBondForce between the displaced ligand and the host): the
    fused Impl evaluates it at x and x + d on the host and feeds force_state{1,2}_ext / energy_ext through
    atm_host_pipeline_step; the reference orchestration clones it into the two inner contexts (ref:
    openmmapi/src/ATMMetaForceImpl.cpp:51-65,113-116).  Both must give U1 + E_bond(x), U2 + E_bond(x + d), the same
    u_sc / energy / forces; at lambda1 == lambda2 (sp = 1/2 below the soft-core threshold) the force changes by exactly
    (F_bond(x) + F_bond(x + d)) / 2."""
    from atmmetaforce import _atmmetaforce_core as core, _capi
    from helpers import rel_rms
    if "CUDA" not in core.getPlatformNames():
        core.registerCudaPlatform()
    core.loadPluginLibrary(PLUGIN)
    kcal = 4.184
    pos, displ = abfe["pos"], abfe["displ"]
    for params, k_bond in (((0.5, 0.5, 0.0, 0.0, 0.0, 200.0 * kcal, 100.0 * kcal, 0.0625, 1.0), 2.0),
                           ((0.2, 0.7, 0.05, 30.0, 1.0, 200.0 * kcal, 100.0 * kcal, 0.0625, 1.0), 25.0)):
        s0, fc0, _ = _bonded_system(core, abfe, params, None, (1,))
        h0 = core.Context(s0)
        h0.setPositions(pos)
        _, f0 = h0.calcForcesAndEnergy(True, True, 1 << 2)
        rec0 = np.array(h0.getEnergyRecord(fc0))
        s, fc, bonds = _bonded_system(core, abfe, params, k_bond, (1, 3))
        e1b, f1b = _bonds_numpy(pos, bonds)
        e2b, f2b = _bonds_numpy(pos + displ, bonds)
        assert e2b - e1b > 10.0                                        # the bonds do see the displacement
        host = core.Context(s)
        host.setPositions(pos)
        e_h, f_h = host.calcForcesAndEnergy(True, True, 1 << 2)
        rec = np.array(host.getEnergyRecord(fc))
        assert abs(rec[_capi.E_U1] - rec0[_capi.E_U1] - e1b) <= 1e-9 * abs(rec0[_capi.E_U1]) + 1e-9
        assert abs(rec[_capi.E_U2] - rec0[_capi.E_U2] - e2b) <= 1e-9 * abs(rec0[_capi.E_U2]) + 1e-9
        assert abs(rec[_capi.E_U] - rec0[_capi.E_U] - (e2b - e1b)) <= 1e-9 * abs(rec0[_capi.E_U]) + 1e-7
        sc = atm_softcore(params, rec[_capi.E_U])
        assert abs(rec[_capi.E_USC] - sc["u_sc"]) <= 1e-9 * abs(sc["u_sc"]) and abs(rec[_capi.E_SP] - sc["sp"]) <= 1e-12
        if params[0] == params[1]:
            assert rec[_capi.E_SP] == 0.5 == rec0[_capi.E_SP]
            assert np.abs((f_h - f0) - 0.5 * (f1b + f2b)).max() <= 1e-6
        else:
            assert abs(rec[_capi.E_SP] - rec0[_capi.E_SP]) > 1e-3     # the bond energies moved the soft-plus weight
        # only group 3: the bond force on its own, through its own Impl
        e3, f3 = host.calcForcesAndEnergy(True, True, 1 << 3)
        assert abs(e3 - e1b) <= 1e-12 * abs(e1b) and np.allclose(f3, f1b, rtol=1e-12, atol=1e-9)
        # the reference orchestration: two inner contexts evaluate NonbondedForce + HarmonicBondForce
        cu = core.Context(s, "CUDA", {"Precision": "mixed"})
        assert cu.usesPlatformKernel(fc)
        cu.setPositions(pos)
        e_k, f_k = cu.calcForcesAndEnergy(True, True, 1 << 2)
        rec_k = np.array(cu.getEnergyRecord(fc))
        assert abs(rec_k[_capi.E_U1] - rec[_capi.E_U1]) <= 1e-6 * abs(rec[_capi.E_U1])
        assert abs(rec_k[_capi.E_USC] - rec[_capi.E_USC]) <= 5e-3
        assert abs(e_k - e_h) <= 1e-5 * abs(e_h) + 1e-3
        assert rel_rms(f_k, f_h) <= 1e-5


def atm_softcore(params, u):
    import atmmetaforce as atm
    return atm.softcore_softplus(params, 0.0, u if params[8] > 0 else -u)


def test_example_scripts_agree_across_the_three_paths():
    """example/abfe and example/rbfe single-point scripts (ref: example/abfe/abfe.py, example/rbfe/rbfe.py): the Python
    Context, the C++ fused Impl and the plugin-glue path print the same sample line (u to 1e-3 kcal/mol), and the ABFE
    one reproduces the reference's pin."""
    import sys
    from atmmetaforce import io

    def sample(script, *flags):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "example", script)] + list(flags), capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        lines = [ln for ln in out.stdout.splitlines() if len(ln.split()) == 9 and ln.split()[0].replace(".", "").isdigit()]
        return io.parse_sample_line(lines[0])

    a = [sample("abfe/abfe_single_point.py"), sample("abfe/abfe_single_point.py", "--cpp")]
    assert abs(a[0]["pert_energy"] - 58.2) <= 0.1 and abs(a[0]["pert_energy"] - a[1]["pert_energy"]) <= 5e-3
    r = [sample("rbfe/rbfe_single_point.py"), sample("rbfe/rbfe_single_point.py", "--cpp"), sample("rbfe/rbfe_single_point.py", "--platform")]
    for x in r[1:]:
        assert abs(x["pert_energy"] - r[0]["pert_energy"]) <= 5e-3
        assert abs(x["pot_energy"] - r[0]["pot_energy"]) <= 1e-6 * abs(r[0]["pot_energy"]) + 0.05
    assert abs(r[0]["pert_energy"] - 2.107) <= 0.1      # survey-time value with the exact Ewald sum (unpinned by the reference)
