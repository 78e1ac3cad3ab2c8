"""atm_host_pipeline_* (the step with pinned HOST buffers on both sides) against the device-buffer path.

Every accumulation of the step is fixed point, so the forces the pipeline returns to the host must be BIT-IDENTICAL to
what atm_step leaves on the device for the same coordinates and the same pair-list history (rebuild / prune / none),
whatever the chunking of the replicas; the energy records likewise.  Parity of atm_step itself against the oracle is
tests/test_gpu_nb2.py; one oracle check is repeated here through the host path.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup(atm, s, params_rows, replicas):
    n = s["pos"].shape[0]
    be = atm.ATMBackend(n, precision="mixed", num_replicas=replicas)
    be.set_displacements(s["displ"])
    be.set_box(s["box"])
    for r in range(replicas):
        be.set_parameters(params_rows[r], replica=r)
    be.nb_setup(s["charge"], s["sigma"], s["epsilon"], s["cutoff"], s["ewald_alpha"], skin=0.1, skin_outer=0.3,
                exclusions=s["excl"], exception_pairs=s["exc14"], exception_params=s["exc14_par"])
    return be


def _coords(s, P, replicas, seed, sigma):
    n = s["pos"].shape[0]
    rng = np.random.default_rng(seed)
    posq = np.zeros((replicas, P, 4), np.float32)
    for r in range(replicas):
        posq[r, :n, :3] = s["pos"] + rng.normal(0, sigma, (n, 3))
        posq[r, :n, 3] = s["charge"]
    return posq


@pytest.mark.parametrize("split", [(3,), (2, 1), (1, 1, 1)])
def test_pipeline_matches_device_path(split):
    import torch
    import atmmetaforce as atm
    from atmmetaforce import synthetic, _capi
    s = synthetic.water_box(12000)   # 4.9 nm box: half the box must hold cutoff + outer skin + a cluster's extent
    sched = synthetic.atm_schedule_22()
    R = sum(split)
    rows = [sched[(5 * r + 3) % 22] for r in range(R)]
    stream = torch.cuda.Stream()
    # device-buffer path: one handle with all R replicas
    ref = _setup(atm, s, rows, R)
    P = ref.P
    # host path: the same replicas split over len(split) handles
    bounds = np.cumsum((0,) + tuple(split))
    bes = [_setup(atm, s, rows[bounds[c]:bounds[c + 1]], split[c]) for c in range(len(split))]
    pipe = atm.HostPipeline(bes)
    posq_h = [torch.zeros((k, P, 4), dtype=torch.float32).pin_memory() for k in split]
    force_h = [torch.full((k, 3 * P), 7, dtype=torch.int64).pin_memory() for k in split]
    en_h = [torch.zeros((k, _capi.NUM_ENERGY_SLOTS), dtype=torch.float64).pin_memory() for k in split]
    C = pipe.PRUNE_CONCURRENT   # the prune runs on the side stream during the step and serves the NEXT step
    plan = [pipe.REBUILD, pipe.NONE, pipe.PRUNE, pipe.NONE, C, pipe.NONE, C, C, pipe.PRUNE, pipe.NONE, pipe.REBUILD, pipe.NONE,
            pipe.PRUNE, pipe.REBUILD, C, pipe.NONE, pipe.REBUILD, pipe.NONE, C, C, C, pipe.REBUILD, pipe.NONE, pipe.PRUNE]
    for it, maint in enumerate(plan):
        x = _coords(s, P, R, seed=100 + it, sigma=0.002 * (1 + it % 3))
        xd = torch.from_numpy(x).cuda()
        force = torch.zeros((R, 3 * P), dtype=torch.int64, device="cuda")
        with torch.cuda.stream(stream):
            if maint == pipe.REBUILD:
                ref.rebuild(xd, stream=stream)
            elif maint == pipe.PRUNE:
                ref.prune(xd, stream=stream)
            ref.step(xd, force, include_energy=True, graph=(it % 2 == 1), stream=stream, concurrent_prune=(maint == C))
        en_ref = ref.get_energies(stream=stream)
        for c in range(len(split)):
            posq_h[c].copy_(torch.from_numpy(x[bounds[c]:bounds[c + 1]]))
        pipe.step(posq_h, force_h, en_h, maintenance=maint, stream=stream)
        stream.synchronize()
        got_f = torch.cat(force_h, 0)
        got_e = torch.cat(en_h, 0).numpy()
        assert torch.equal(got_f, force.cpu()), f"step {it} (maintenance {maint}): forces differ"
        # U1, U2, u, u_sc, ebias, energy, sp are deterministic (fixed-point partial sums)
        assert np.array_equal(got_e[:, :7], en_ref[:, :7]), f"step {it}: energy records differ"
        assert np.abs(got_f.numpy()).max() > 0
    # launches are accounted per handle: pack + nb2 + merge per plain step at least
    assert all(b.launch_count() >= 3 * len(plan) for b in bes)
    pipe.close()
    for b in bes + [ref]:
        b.close()


def test_pipeline_against_oracle(abfe):
    """The reference fixture through the host path: U1, u and the merged force against the CPU oracle."""
    import torch
    import atmmetaforce as atm
    import oracle_py as O
    from atmmetaforce import _capi
    from helpers import oracle_system, force_from_fixed, rel_rms
    s = dict(abfe)
    s["cutoff"], s["ewald_alpha"] = 1.0, O.ewald_alpha(1.0)
    n = s["pos"].shape[0]
    be = _setup(atm, s, [s["params"]], 1)
    P = be.P
    pipe = atm.HostPipeline([be])
    posq_h = torch.zeros((1, P, 4), dtype=torch.float32).pin_memory()
    posq_h[0, :n, :3] = torch.from_numpy(s["pos"].astype(np.float32))
    posq_h[0, :n, 3] = torch.from_numpy(s["charge"].astype(np.float32))
    force_h = torch.zeros((1, 3 * P), dtype=torch.int64).pin_memory()
    en_h = torch.zeros((1, _capi.NUM_ENERGY_SLOTS), dtype=torch.float64).pin_memory()
    stream = torch.cuda.Stream()
    pipe.step([posq_h], [force_h], [en_h], maintenance=pipe.REBUILD, stream=stream)
    stream.synchronize()
    pos32 = s["pos"].astype(np.float32).astype(np.float64)
    pos2_32 = (s["pos"].astype(np.float32) + s["displ"].astype(np.float32)).astype(np.float64)
    S = oracle_system(O, s, 1.0, s["ewald_alpha"])
    e1, _, f1 = S.nb_direct(pos32)
    e2, _, f2 = S.nb_direct(pos2_32)
    sc = O.scalars(s["params"], e1, e2)
    f_ref = O.merge_ref(np.zeros_like(f1), f1, f2, sc["sp_ref"], s["params"][8])
    en = en_h.numpy()[0]
    assert abs(en[_capi.E_U1] - e1) <= 1e-6 * abs(e1)
    assert abs(en[_capi.E_USC] - sc["u_sc"]) <= 5e-3
    assert rel_rms(force_from_fixed(force_h.numpy()[0], n, P), f_ref) <= 1e-5
    pipe.close()
    be.close()


def test_force_formats():
    """ATM_FORCE_F32: exactly the fixed-point force converted to float32 ((float)(f / 2^32)); ATM_FORCE_NONE: energies only."""
    import torch
    import atmmetaforce as atm
    from atmmetaforce import synthetic, _capi
    s = synthetic.water_box(12000)
    sched = synthetic.atm_schedule_22()
    be = _setup(atm, s, [sched[4], sched[17]], 2)
    P = be.P
    pipe = atm.HostPipeline([be])
    posq_h = torch.from_numpy(_coords(s, P, 2, seed=3, sigma=0.003)).pin_memory()
    f64 = torch.zeros((2, 3 * P), dtype=torch.int64).pin_memory()
    f32 = torch.zeros((2, 3 * P), dtype=torch.float32).pin_memory()
    en_a, en_b, en_c = (torch.zeros((2, _capi.NUM_ENERGY_SLOTS), dtype=torch.float64).pin_memory() for _ in range(3))
    stream = torch.cuda.Stream()
    pipe.step([posq_h], [f64], [en_a], maintenance=pipe.REBUILD, stream=stream)
    stream.synchronize()
    pipe.step([posq_h], [f32], [en_b], maintenance=pipe.NONE, stream=stream)
    stream.synchronize()
    pipe.step([posq_h], None, [en_c], maintenance=pipe.NONE, stream=stream)
    stream.synchronize()
    expect = (f64.numpy().astype(np.float64) / 4294967296.0).astype(np.float32)
    assert np.array_equal(f32.numpy(), expect) and np.abs(expect).max() > 0
    assert np.array_equal(en_a.numpy()[:, :7], en_b.numpy()[:, :7]) and np.array_equal(en_a.numpy()[:, :7], en_c.numpy()[:, :7])
    # ATM_POSQ_F3: packed float3 coordinates give the same bits as the float4 posq (the charges come from nb_setup)
    pos3 = posq_h[:, :, :3].contiguous().pin_memory()
    f64b = torch.zeros_like(f64).pin_memory()
    pipe.step([pos3], [f64b], [en_b], maintenance=pipe.PRUNE, stream=stream)
    stream.synchronize()
    pipe.step([posq_h], [f64], [en_a], maintenance=pipe.PRUNE, stream=stream)
    stream.synchronize()
    assert torch.equal(f64, f64b) and np.array_equal(en_a.numpy()[:, :7], en_b.numpy()[:, :7])
    with pytest.raises(atm.ATMError, match="nothing to return"):
        pipe.step([posq_h], None, None, maintenance=pipe.NONE, stream=stream)
    with pytest.raises(atm.ATMError, match="int64"):
        pipe.step([posq_h], [torch.zeros((2, 3 * P), dtype=torch.float64).pin_memory()], [en_a], maintenance=pipe.NONE, stream=stream)
    pipe.close()
    be.close()


def test_pipeline_errors():
    import torch
    import atmmetaforce as atm
    from atmmetaforce import synthetic
    s = synthetic.water_box(12000)
    be = _setup(atm, s, [synthetic.atm_schedule_22()[0]], 1)
    P = be.P
    with pytest.raises(atm.ATMError):
        atm.HostPipeline([be, be])                      # the same handle twice
    pipe = atm.HostPipeline([be])
    posq_h = torch.zeros((1, P, 4), dtype=torch.float32).pin_memory()
    force_h = torch.zeros((1, 3 * P), dtype=torch.int64).pin_memory()
    stream = torch.cuda.Stream()
    with pytest.raises(atm.ATMError, match="rebuild"):  # no pair list yet and none asked for
        pipe.step([posq_h], [force_h], maintenance=pipe.NONE, stream=stream)
    with pytest.raises(atm.ATMError, match="pinned"):
        pipe.step([torch.zeros((1, P, 4))], [force_h], maintenance=pipe.REBUILD, stream=stream)
    with pytest.raises(atm.ATMError):
        pipe.step([posq_h], [force_h], maintenance=5, stream=stream)
    pipe.close()
    be.close()


def test_external_state_terms_through_the_host_path():
    """atm_host_io.force_state{1,2}_ext_host / energy_ext_host (the generic hook for other variable-group forces evaluated by
    the caller): bit-identical to atm_step with the same arrays as device buffers, over two chunks; and the merged force
    differs from the run without them by exactly sp * F2_ext + (1 - sp) * F1_ext at equal sp."""
    import torch
    import atmmetaforce as atm
    from atmmetaforce import synthetic, _capi
    s = synthetic.water_box(12000)
    sched = synthetic.atm_schedule_22()
    rows = [sched[4], sched[11], sched[17]]
    ref = _setup(atm, s, rows, 3)
    P, n = ref.P, s["pos"].shape[0]
    bes = [_setup(atm, s, rows[:2], 2), _setup(atm, s, rows[2:], 1)]
    pipe = atm.HostPipeline(bes)
    rng = np.random.default_rng(8)
    x = _coords(s, P, 3, seed=21, sigma=0.003)
    ext = np.zeros((2, 3, 3, P))
    ext[:, :, :, :n] = rng.normal(0, 300.0, (2, 3, 3, n))
    ext_fixed = np.rint(ext * 4294967296.0).astype(np.int64).reshape(2, 3, 3 * P)
    e_ext = rng.normal(0, 20.0, (3, 2))
    stream = torch.cuda.Stream()
    xd = torch.from_numpy(x).cuda()
    force = torch.zeros((3, 3 * P), dtype=torch.int64, device="cuda")
    with torch.cuda.stream(stream):
        ref.rebuild(xd, stream=stream)
        ref.step(xd, force, f1_ext=torch.from_numpy(ext_fixed[0]).cuda(), f2_ext=torch.from_numpy(ext_fixed[1]).cuda(),
                 energy_ext=torch.from_numpy(e_ext).cuda(), include_energy=True, stream=stream)
    en_ref = ref.get_energies(stream=stream)
    split = (slice(0, 2), slice(2, 3))
    posq_h = [torch.from_numpy(x[c]).pin_memory() for c in split]
    force_h = [torch.zeros((k, 3 * P), dtype=torch.int64).pin_memory() for k in (2, 1)]
    en_h = [torch.zeros((k, _capi.NUM_ENERGY_SLOTS), dtype=torch.float64).pin_memory() for k in (2, 1)]
    f1_h = [torch.from_numpy(ext_fixed[0][c].copy()).pin_memory() for c in split]
    f2_h = [torch.from_numpy(ext_fixed[1][c].copy()).pin_memory() for c in split]
    ee_h = [torch.from_numpy(e_ext[c].copy()).pin_memory() for c in split]
    pipe.step(posq_h, force_h, en_h, maintenance=pipe.REBUILD, stream=stream, f1_ext_host=f1_h, f2_ext_host=f2_h, energy_ext_host=ee_h)
    stream.synchronize()
    pipe.step(posq_h, force_h, en_h, maintenance=pipe.NONE, stream=stream, f1_ext_host=f1_h, f2_ext_host=f2_h, energy_ext_host=ee_h)
    stream.synchronize()
    assert torch.equal(torch.cat(force_h, 0), force.cpu())
    got_e = torch.cat(en_h, 0).numpy()
    assert np.array_equal(got_e[:, :7], en_ref[:, :7])
    # the same step without the external energies but with the external forces: sp is unchanged by the forces, so the
    # difference to a plain step is the sp-weighted mix of the two external force sets
    plain_f = [torch.zeros((k, 3 * P), dtype=torch.int64).pin_memory() for k in (2, 1)]
    pipe.step(posq_h, plain_f, en_h, maintenance=pipe.NONE, stream=stream)
    stream.synchronize()
    sp = torch.cat(en_h, 0).numpy()[:, _capi.E_SP]
    only_f = [torch.zeros((k, 3 * P), dtype=torch.int64).pin_memory() for k in (2, 1)]
    pipe.step(posq_h, only_f, en_h, maintenance=pipe.NONE, stream=stream, f1_ext_host=f1_h, f2_ext_host=f2_h)
    stream.synchronize()
    assert np.array_equal(torch.cat(en_h, 0).numpy()[:, _capi.E_SP], sp)
    diff = (torch.cat(only_f, 0) - torch.cat(plain_f, 0)).numpy() / 4294967296.0
    want = sp[:, None] * ext[1].reshape(3, 3 * P) + (1.0 - sp[:, None]) * ext[0].reshape(3, 3 * P)
    assert np.abs(diff - want).max() <= 1e-6
    # argument checks
    with pytest.raises(atm.ATMError, match="wrong dtype or size"):
        pipe.step(posq_h, force_h, en_h, stream=stream, energy_ext_host=[torch.zeros(3).pin_memory(), torch.zeros(2).pin_memory()])
    with pytest.raises(atm.ATMError, match="pinned"):
        pipe.step(posq_h, force_h, en_h, stream=stream, energy_ext_host=[torch.zeros(4, dtype=torch.float64), torch.zeros(2, dtype=torch.float64)])
    pipe.close()
    for b in bes + [ref]:
        b.close()


def test_pipeline_with_pme():
    """The host path with the two-state PME inside the step: one chunk holding both replicas is bit-identical to atm_step
    on device buffers (same kernels, same transform batch); two one-replica chunks agree to float-mesh accuracy (their
    transforms are batched differently)."""
    import torch
    import atmmetaforce as atm
    from atmmetaforce import synthetic, _capi
    s = synthetic.water_box(12000)
    sched = synthetic.atm_schedule_22()
    rows = [sched[6], sched[15]]
    grid = synthetic.pme_grid(s["box"], s["ewald_alpha"])

    def setup(rr):
        be = _setup(atm, s, rr, len(rr))
        be.pme_setup(grid)
        return be

    ref = setup(rows)
    P = ref.P
    x = _coords(s, P, 2, seed=11, sigma=0.003)
    xd = torch.from_numpy(x).cuda()
    force = torch.zeros((2, 3 * P), dtype=torch.int64, device="cuda")
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        ref.rebuild(xd, stream=stream)
        ref.step(xd, force, include_energy=True, stream=stream)
    en_ref = ref.get_energies(stream=stream)
    for split in ((2,), (1, 1)):
        bounds = np.cumsum((0,) + split)
        bes = [setup(rows[bounds[c]:bounds[c + 1]]) for c in range(len(split))]
        pipe = atm.HostPipeline(bes)
        posq_h = [torch.from_numpy(x[bounds[c]:bounds[c + 1]]).pin_memory() for c in range(len(split))]
        force_h = [torch.zeros((k, 3 * P), dtype=torch.int64).pin_memory() for k in split]
        en_h = [torch.zeros((k, _capi.NUM_ENERGY_SLOTS), dtype=torch.float64).pin_memory() for k in split]
        pipe.step(posq_h, force_h, en_h, maintenance=pipe.REBUILD, stream=stream)
        stream.synchronize()
        pipe.check()
        pipe.step(posq_h, force_h, en_h, maintenance=pipe.NONE, stream=stream)
        stream.synchronize()
        got_f, got_e = torch.cat(force_h, 0), torch.cat(en_h, 0).numpy()
        if split == (2,):
            assert torch.equal(got_f, force.cpu())
            assert np.array_equal(got_e[:, :14], en_ref[:, :14])
        else:
            a, b = got_f.numpy().astype(np.float64), force.cpu().numpy().astype(np.float64)
            assert np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum()) <= 2e-6
            assert np.abs(got_e[:, _capi.E_USC] - en_ref[:, _capi.E_USC]).max() <= 1e-3
            assert np.abs(got_e[:, 11] / en_ref[:, 11] - 1.0).max() <= 1e-6
        pipe.close()
        for b in bes:
            b.close()
    ref.close()
