"""GPU parity of the Tier-1 kernels (CopyState, HybridForce, execute) against the CPU oracle -- bit exact.

Reference semantics: platforms/common/src/kernels/atmmetaforce.cc:1-52; the reference itself has no kernel unit
tests, so the oracle (oracle/atm_oracle.c) is the checker.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32 if a.dtype == np.float32 else np.uint64)


@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 1000, 8479, 21559, 100003])
@pytest.mark.parametrize("mode", ["single", "mixed", "double"])
def test_copy_state_bit_exact(n, mode):
    import torch
    import atmmetaforce as atm
    import oracle_py as O
    rng = np.random.default_rng(7 + n)
    P = 32 * ((n + 31) // 32) if n else 32
    be = atm.ATMBackend(n, padded_num_particles=P, precision=mode)
    perm = rng.permutation(n).astype(np.int32)
    dxyz = np.zeros((n, 3))
    nl = min(n, 21)
    lig = rng.choice(n, nl, replace=False) if n else np.zeros(0, int)
    dxyz[lig] = [2.2, -2.2, 0.3]
    if n > 40:
        dxyz[rng.choice(n, 17, replace=False)] = [-2.2, -2.2, -2.2]
    be.set_displacements(dxyz, atom_index=perm)
    table = O.displ_table(n, P, perm, dxyz)
    real = np.float64 if mode == "double" else np.float32
    posq = np.zeros((P, 4), real)
    posq[:n] = rng.uniform(-6, 10, (n, 4))
    posq[:n:7, 3] = -0.0  # exercise the "+ 0" on the charge slot
    corr = rng.uniform(-1e-7, 1e-7, (P, 4)).astype(np.float32) if mode == "mixed" else None
    sentinel = 12345.0
    d_posq = torch.from_numpy(posq).cuda()
    d_p1 = torch.full((P, 4), sentinel, dtype=d_posq.dtype, device="cuda")
    d_p2 = torch.full((P, 4), sentinel, dtype=d_posq.dtype, device="cuda")
    if mode == "mixed":
        d_c = torch.from_numpy(corr).cuda()
        d_c1, d_c2 = torch.full_like(d_c, sentinel), torch.full_like(d_c, sentinel)
        be.copy_state(d_posq, d_p1, d_p2, d_c, d_c1, d_c2)
    else:
        be.copy_state(d_posq, d_p1, d_p2)
    torch.cuda.synchronize()
    if mode == "double":
        e1, e2 = O.copy_state_f64(posq[:n], table[:n])
    else:
        e1, _, e2, _ = O.copy_state_f32(posq[:n], corr[:n] if corr is not None else None, table[:n])
    p1, p2 = d_p1.cpu().numpy(), d_p2.cpu().numpy()
    assert np.array_equal(_bits(p1[:n]), _bits(e1))
    assert np.array_equal(_bits(p2[:n]), _bits(e2))
    # padded tail untouched
    assert np.all(p1[n:] == sentinel) and np.all(p2[n:] == sentinel)
    if mode == "mixed":
        assert np.array_equal(_bits(d_c1.cpu().numpy()[:n]), _bits(corr[:n]))
        assert np.array_equal(_bits(d_c2.cpu().numpy()[:n]), _bits(corr[:n]))
    # idempotence / linearity property at any size: posq2 - posq1 == float(d) wherever no rounding occurs
    be.close()


def test_copy_state_requires_corrections_in_mixed():
    import torch
    import atmmetaforce as atm
    be = atm.ATMBackend(64, precision="mixed")
    be.set_displacements(np.zeros((64, 3)))
    t = torch.zeros((64, 4), device="cuda")
    with pytest.raises(atm.ATMError):
        be.copy_state(t, t.clone(), t.clone())
    be.close()


@pytest.mark.parametrize("n,P", [(1, 32), (2, 32), (33, 64), (8479, 8480), (21559, 21568), (5, 7), (100001, 100032)])
@pytest.mark.parametrize("sp", [0.0, 1.0, 0.5, 0.123456789, -0.25, 1.75])
def test_hybrid_force_bit_exact(n, P, sp):
    import torch
    import atmmetaforce as atm
    import oracle_py as O
    rng = np.random.default_rng(n + int(sp * 1000) % 97)
    be = atm.ATMBackend(n, padded_num_particles=P, precision="mixed")

    def rand_forces():
        f = rng.normal(0, 2000.0, 3 * P) * 2.0 ** 32
        f[rng.integers(0, 3 * P, 5)] *= 1e3  # a few very large entries
        return f.astype(np.int64)
    f0, f1, f2 = rand_forces(), rand_forces(), rand_forces()
    d0 = torch.from_numpy(f0.copy()).cuda()
    be.hybrid_force(d0, torch.from_numpy(f1).cuda(), torch.from_numpy(f2).cuda(), sp)
    torch.cuda.synchronize()
    got = d0.cpu().numpy()
    exp = O.hybrid_force_i64(n, P, f0, f1, f2, sp)
    assert np.array_equal(got, exp)
    # padded tail untouched
    for c in range(3):
        assert np.array_equal(got[c * P + n:(c + 1) * P], f0[c * P + n:(c + 1) * P])
    be.close()


def test_hybrid_force_close_to_reference_platform_double():
    """<= 1e-5 relative RMS against the Reference-platform merge in double (ReferenceATMMetaForceKernels.cpp:101-110)."""
    import torch
    import atmmetaforce as atm
    import oracle_py as O
    from helpers import force_from_fixed, rel_rms, FIX
    rng = np.random.default_rng(3)
    n, P = 8479, 8480
    f1 = rng.normal(0, 1000, (n, 3))
    f2 = f1 + rng.normal(0, 50, (n, 3))
    base = rng.normal(0, 500, (n, 3))

    def to_fixed(f):
        b = np.zeros(3 * P, np.int64)
        b.reshape(3, P)[:, :n] = np.rint(f.T * FIX).astype(np.int64)
        return b
    for direction, params in ((1.0, [0.3, 0.6, 0.05, 100.0, 0, 800, 400, 0.0625, 1.0]),
                              (-1.0, [0.3, 0.6, 0.05, 100.0, 0, 800, 400, 0.0625, -1.0])):
        be = atm.ATMBackend(n, padded_num_particles=P)
        be.set_parameters(params)
        U1, U2 = -1000.0, -880.0
        d0 = torch.from_numpy(to_fixed(base)).cuda()
        energy = be.execute(U1, U2, d0, torch.from_numpy(to_fixed(f1)).cuda(), torch.from_numpy(to_fixed(f2)).cuda())
        torch.cuda.synchronize()
        sc = O.scalars(params, U1, U2)
        exp = O.merge_ref(base, f1, f2, sc["sp_ref"], direction)
        got = force_from_fixed(d0.cpu().numpy(), n, P)
        assert rel_rms(got, exp) < 1e-8
        assert abs(energy - sc["energy"]) <= 1e-12 * abs(sc["energy"])
        assert abs(be.get_perturbation_energy() - sc["u_sc"]) <= 1e-12 * abs(sc["u_sc"])
        be.close()


def test_wrap_positions():
    import torch
    import atmmetaforce as atm
    rng = np.random.default_rng(5)
    n = 1000
    be = atm.ATMBackend(n)
    posq = rng.uniform(-12, 15, (be.P, 4)).astype(np.float32)
    box = np.diag([4.2, 4.6, 4.3])
    out = torch.zeros((be.P, 4), device="cuda")
    be.wrap_positions(torch.from_numpy(posq).cuda(), out, box)
    w = out.cpu().numpy()[:n]
    L = np.diag(box).astype(np.float32)
    assert np.all(w[:, :3] >= -1e-5) and np.all(w[:, :3] <= L + 1e-5)
    delta = (w[:, :3] - posq[:n, :3]) / L
    assert np.allclose(delta, np.rint(delta), atol=1e-4)
    assert np.array_equal(w[:, 3], posq[:n, 3])
    be.close()
