"""Property-based CPU tests (hypothesis) of the host side: XML round trips of arbitrary forces, the host scalar stage
against the oracle over the whole parameter space, the replica-exchange sweep, the stand-in HarmonicBondForce, and the
bench's chunk split.  No GPU needed."""
import importlib.util
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import atmmetaforce as atm
import oracle_py as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
finite = dict(allow_nan=False, allow_infinity=False)


@settings(max_examples=60, deadline=None)
@given(params=st.tuples(st.floats(0, 1, **finite), st.floats(0, 1, **finite), st.floats(0, 0.5, **finite), st.floats(-500, 500, **finite),
                        st.floats(-50, 50, **finite), st.floats(1, 1e4, **finite), st.floats(0.5, 5e3, **finite), st.floats(0.01, 1, **finite),
                        st.sampled_from([1.0, -1.0])),
       particles=st.lists(st.tuples(st.floats(-50, 50, **finite), st.floats(-50, 50, **finite), st.floats(-50, 50, **finite)), max_size=12),
       groups=st.lists(st.integers(0, 31), max_size=4, unique=True), group=st.integers(0, 31), name=st.text("abcXYZ_ 09", max_size=8))
def test_xml_round_trip_of_arbitrary_forces(params, particles, groups, group, name):
    """serialize -> deserialize -> serialize is the identity on the document, and every number survives exactly
    (ref schema: serialization/src/ATMMetaForceProxy.cpp:17-55)."""
    f = atm.ATMMetaForce(*params, groups)
    f.setForceGroup(group)
    f.setName(name)
    for i, d in enumerate(particles):
        f.addParticle(i, *d)
    xml = atm.serialize(f)
    g = atm.deserialize(xml)
    assert atm.serialize(g) == xml
    assert g.getNumParticles() == len(particles) and g.getForceGroup() == group and list(g.getVariableForceGroups()) == groups
    for i, d in enumerate(particles):
        assert tuple(g.getParticleParameters(i)) == (i,) + tuple(d)
    assert (g.getDefaultLambda1(), g.getDefaultUmax(), g.getDefaultDirection()) == (params[0], params[5], params[8])


@settings(max_examples=300, deadline=None)
@given(lam1=st.floats(0, 1, **finite), lam2=st.floats(0, 1, **finite), alpha=st.one_of(st.just(0.0), st.floats(1e-4, 0.5, **finite)),
       u0=st.floats(-100, 600, **finite), w0=st.floats(-10, 10, **finite), umax=st.floats(50, 2000, **finite), frac=st.floats(0.05, 0.95, **finite),
       acore=st.floats(0.01, 1, **finite), direction=st.sampled_from([1.0, -1.0]), U1=st.floats(-3e5, 1e3, **finite), du=st.floats(-1e3, 1e5, **finite))
def test_host_scalar_stage_matches_oracle_everywhere(lam1, lam2, alpha, u0, w0, umax, frac, acore, direction, U1, du):
    """atm_softcore_softplus (the arithmetic the device scalar stage runs) against the oracle's restatement of
    CommonATMMetaForceKernels.cpp:19-30,182-199 over all three soft-core branches and both directions."""
    p = [lam1, lam2, alpha, u0, w0, umax, frac * umax, acore, direction]
    a = atm.softcore_softplus(p, U1, U1 + du)
    b = O.scalars(p, U1, U1 + du)
    for k in ("u_sc", "fp", "ebias", "bfp", "energy", "sp", "sp_ref"):
        if np.isinf(b[k]) or np.isnan(b[k]):
            assert (np.isinf(a[k]) and a[k] == b[k]) or (np.isnan(a[k]) and np.isnan(b[k]))
        else:
            assert abs(a[k] - b[k]) <= 1e-11 * max(1.0, abs(b[k])), (k, a[k], b[k])
    assert 0.0 <= a["fp"] <= 1.0 + 1e-15


@settings(max_examples=80, deadline=None)
@given(n=st.integers(2, 24), seed=st.integers(0, 2 ** 31 - 1), cycle=st.integers(0, 10 ** 6), data=st.data())
def test_hrex_sweep_keeps_a_permutation_and_is_deterministic(n, seed, cycle, data):
    sched = np.array([[k / max(n - 1, 1), k / max(n - 1, 1), 0.0, 0.0, 0.0, 836.8, 418.4, 0.0625, 1.0] for k in range(n)])
    u12 = np.array(data.draw(st.lists(st.tuples(st.floats(-1e5, 0, **finite), st.floats(-1e5, 1e3, **finite)), min_size=n, max_size=n)))
    state0 = np.array(data.draw(st.permutations(list(range(n)))), dtype=np.int32)
    beta = 1.0 / (0.0083144626 * 300.0)
    s1, acc1 = atm.hrex_sweep(sched, u12, state0, beta, seed, cycle)
    s2, acc2 = atm.hrex_sweep(sched, u12, state0, beta, seed, cycle)
    assert sorted(s1) == list(range(n)) and np.array_equal(s1, s2) and acc1 == acc2
    assert np.array_equal(state0, np.array(state0))            # the input is not modified


@settings(max_examples=60, deadline=None)
@given(seed=st.integers(0, 10 ** 6), periodic=st.booleans(), shift=st.tuples(st.floats(-3, 3, **finite), st.floats(-3, 3, **finite), st.floats(-3, 3, **finite)))
def test_standin_harmonic_bond_force_properties(seed, periodic, shift):
    """The stand-in HarmonicBondForce (what the fused Impl feeds through the external per-state inputs): forces are minus
    the finite-difference gradient, sum to zero, and a rigid translation changes nothing."""
    from atmmetaforce import _atmmetaforce_core as core
    rng = np.random.default_rng(seed)
    n, nb = 12, 7
    pos = rng.uniform(0.2, 2.8, (n, 3))
    a, b = rng.integers(0, n // 2, nb), rng.integers(n // 2, n, nb)
    s = core.System()
    for _ in range(n):
        s.addParticle(1.0)
    s.setDefaultPeriodicBoxVectors([3.0, 0, 0], [0, 3.0, 0], [0, 0, 3.0])
    s.addHarmonicBondForce(a.tolist(), b.tolist(), rng.uniform(0.1, 0.4, nb).tolist(), rng.uniform(10, 1e4, nb).tolist(),
                           forceGroup=0, usesPeriodicBoundaryConditions=periodic)
    ctx = core.Context(s)

    def ef(x):
        ctx.setPositions(x)
        return ctx.calcForcesAndEnergy(True, True, 1)

    e, f = ef(pos)
    assert np.abs(f.sum(0)).max() <= 1e-9 * (1.0 + np.abs(f).max())
    e_t, f_t = ef(pos + np.array(shift))
    if periodic or True:       # relative vectors are unchanged by a rigid shift with or without the minimum image
        assert abs(e_t - e) <= 1e-9 * (1.0 + abs(e)) and np.allclose(f_t, f, rtol=1e-8, atol=1e-7 * (1.0 + np.abs(f).max()))
    i, c, h = int(rng.integers(0, n)), int(rng.integers(0, 3)), 1e-6
    xp, xm = pos.copy(), pos.copy()
    xp[i, c] += h
    xm[i, c] -= h
    fd = -(ef(xp)[0] - ef(xm)[0]) / (2 * h)
    assert abs(fd - f[i, c]) <= 1e-4 * (1.0 + np.abs(f).max())


@given(R=st.integers(1, 64), k=st.integers(1, 12), mode=st.sampled_from(["equal", "graded", "auto"]))
def test_bench_chunk_split_partitions_the_replicas(R, k, mode):
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    sizes = bench.e2e_split(mode, R, k)
    assert sum(sizes) == R and min(sizes) >= 1
    with pytest.raises(SystemExit):
        bench.e2e_split("1,1", R + 5, k)
