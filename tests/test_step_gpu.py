"""GPU parity of cavb200_step, the one-launch cavity force + Bussi step, for every kernel variant
(0 = reduce kernel + apply kernel, 1 = fused persistent kernel with one hand-off, 2 = split-phase
persistent kernel, 3 = split-phase with a folder CTA, the default) on the edge cases of the reference: photon first / middle / last /
absent / duplicated (src/CavityForceCompute.cc:73-89,120-126,149-156), thermostat groups that are a
prefix, a window or empty, dt = 0 (src/BussiReservoirThermostat.h:45-48), zero kinetic energy
(:57-61), sizes below one warp, ragged sizes around the CTA size, launch shapes.

Tolerances: forces / energies 1e-10 relative (BASELINE), velocities and alpha 1e-12 relative, photon
index, zero components and untouched velocities exact; variants 0-2 must agree BIT FOR BIT (same
per-thread summation orders); variant 3 streams with one CTA fewer, so at full grids its compensated
dipole sum runs over a different partition of the same terms: equal to <= 1e-14 relative (measured:
bit-identical in every case here), bit for bit whenever the grid is not full."""
import numpy as np
import pytest

from cav_hoomd_b200 import capi, synth

pytestmark = pytest.mark.gpu

KT, TAU, DT = synth.KT_100K, synth.TAU_5PS, synth.DT_1FS
VARIANTS = [0, 1, 2, 3]


def run_step(handle, s, first, n, a, omegac=0.01, g=1e-3):
    p = capi.Params.make(omegac, g)
    dev = {k: capi.DeviceArray.from_numpy(getattr(s, k)) for k in ("pos", "charge", "image", "vel")}
    d_f = capi.DeviceArray.from_numpy(np.full((s.N, 4), np.nan))
    handle.bussi_reset()
    handle.step(dev["pos"], dev["charge"], dev["image"], d_f, dev["vel"], s.N, s.box, s.L_typeid, p, first, n, a)
    en, dip, ph = handle.force_read()
    return d_f.numpy(), dev["vel"].numpy(), en, dip, ph, handle.bussi_read()


def check_against_oracle(coracle, s, first, n, a, out, omegac=0.01, g=1e-3):
    f, v, en, dip, ph, bo = out
    ref = coracle.cavity_force(s.pos, s.charge, s.image, s.box, s.L_typeid, omegac, g)
    assert ph == ref["photon_idx"]
    scale = max(np.abs(ref["force"]).max(), 1e-300)
    assert np.abs(f - ref["force"]).max() <= 1e-10 * scale
    assert np.array_equal(f[:, 2] == 0.0, ref["force"][:, 2] == 0.0) and np.all(f[:, 3] == 0.0)
    assert np.allclose(en, ref["energies"], rtol=1e-10, atol=0)
    vref = s.vel.copy()
    if n > 0 and a.deltaT != 0.0:
        idx = np.arange(first, first + n, dtype=np.uint32)
        alpha_ref, ke_ref = coracle.bussi_step(vref, idx, a.dof, a.deltaT, a.kT, a.tau, a.r_normal, a.gamma_draw, np.zeros(2))
        assert abs(bo["alpha"] - alpha_ref) <= 1e-12 * abs(alpha_ref)
    assert np.allclose(v, vref, rtol=1e-12, atol=0)
    # velocities outside the thermostatted window are untouched, bit for bit
    outside = np.ones(s.N, dtype=bool)
    outside[first:first + n] = False
    assert np.array_equal(v[outside].view(np.uint64), s.vel[outside].view(np.uint64))


def args_for(n_group, dt=DT, **kw):
    dof = max(3.0 * n_group - 3.0, 0.0)
    return capi.BussiArgs(KT, TAU, dt, dof, kw.get("r_normal", 0.3), kw.get("gamma_draw", max(dof - 1.0, 0.0) / 2.0 * 1.001))


@pytest.mark.parametrize("n_mol", [1, 2, 31, 32, 33, 383, 384, 385, 1000, 65537, 113665])
def test_step_sizes_all_variants_bitwise(handle, coracle, n_mol):
    s = synth.make_system(n_mol)
    a = args_for(n_mol)
    outs = []
    for variant in VARIANTS:
        handle.set_tuning(variant=variant, threads=384, ctas_per_sm=2, unroll=2)
        outs.append(run_step(handle, s, 0, n_mol, a))
        check_against_oracle(coracle, s, 0, n_mol, a, outs[-1])
    for variant, o in zip(VARIANTS[1:], outs[1:]):
        if variant == 3 and n_mol + 1 > 295 * 384:
            assert np.abs(o[0] - outs[0][0]).max() <= 1e-14 * np.abs(outs[0][0]).max()
            assert np.allclose(o[2], outs[0][2], rtol=1e-14, atol=0)
        else:
            assert np.array_equal(o[0].view(np.uint64), outs[0][0].view(np.uint64))
            assert np.array_equal(o[2], outs[0][2])
        assert np.array_equal(o[1].view(np.uint64), outs[0][1].view(np.uint64))
        assert o[5]["alpha"] == outs[0][5]["alpha"]


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("photon", ["first", "middle", "last", "absent", "duplicated"])
def test_step_photon_placement(handle, coracle, photon, variant):
    handle.set_tuning(variant=variant, threads=384, ctas_per_sm=2, unroll=2)
    n_mol = 20011
    s = synth.make_system(n_mol, photon=photon)
    # thermostat the molecular particles only when they are a contiguous range; else everything
    first, n = (1, n_mol) if photon == "first" else (0, n_mol if photon in ("last", "absent") else s.N)
    a = args_for(n)
    check_against_oracle(coracle, s, first, n, a, run_step(handle, s, first, n, a))


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("first,n", [(0, 0), (0, 1), (0, 5000), (1234, 5000), (19999, 2), (7, 20001 - 7)])
def test_step_group_windows(handle, coracle, first, n, variant):
    handle.set_tuning(variant=variant, threads=384, ctas_per_sm=2, unroll=2)
    s = synth.make_system(20000)
    a = args_for(n)
    check_against_oracle(coracle, s, first, n, a, run_step(handle, s, first, n, a))


@pytest.mark.parametrize("variant", VARIANTS)
def test_step_dt_zero_and_zero_ke(handle, coracle, variant):
    handle.set_tuning(variant=variant, threads=384, ctas_per_sm=2, unroll=2)
    s = synth.make_system(5000)
    # dt == 0: the thermostat half is a no-op (alpha = 1), the force half still runs
    a0 = args_for(5000, dt=0.0)
    out = run_step(handle, s, 0, 5000, a0)
    check_against_oracle(coracle, s, 0, 5000, a0, out)
    assert np.array_equal(out[1].view(np.uint64), s.vel.view(np.uint64))
    # zero kinetic energy with dof != 0: error flag (the reference throws), velocities untouched, forces correct
    s.vel[:, :3] = 0.0
    a = args_for(5000)
    f, v, en, dip, ph, bo = run_step(handle, s, 0, 5000, a)
    assert bo["err"] == 1.0 and np.all(v[:, :3] == 0.0)
    ref = coracle.cavity_force(s.pos, s.charge, s.image, s.box, s.L_typeid, 0.01, 1e-3)
    assert np.abs(f - ref["force"]).max() <= 1e-10 * np.abs(ref["force"]).max()
    handle.bussi_reset()


@pytest.mark.parametrize("variant", [2, 3])
@pytest.mark.parametrize("threads,ctas,unroll", [(128, 2, 2), (256, 4, 2), (384, 1, 2), (512, 1, 4), (768, 1, 2), (1024, 1, 2), (256, 1, 8)])
def test_step_launch_shapes_split_kernel(handle, coracle, threads, ctas, unroll, variant):
    handle.set_tuning(variant=variant, threads=threads, ctas_per_sm=ctas, unroll=unroll)
    s = synth.make_system(150001)
    a = args_for(150001)
    check_against_oracle(coracle, s, 0, 150001, a, run_step(handle, s, 0, 150001, a))
    handle.set_tuning(variant=3, threads=384, ctas_per_sm=2, unroll=2)


@pytest.mark.parametrize("variant", [2, 3])
def test_step_back_to_back_epochs(handle, coracle, variant):
    """Many launches in a row on alternating systems and grid sizes: the hand-off's epoch tagging must
    never let a record of an earlier launch pass for a current one."""
    handle.set_tuning(variant=variant, threads=384, ctas_per_sm=2, unroll=2)
    systems = [synth.make_system(n, replica=k) for k, n in enumerate((300, 70001, 5, 20000))]
    p = capi.Params.make(0.01, 1e-3)
    devs = []
    for s in systems:
        d = {k: capi.DeviceArray.from_numpy(getattr(s, k)) for k in ("pos", "charge", "image", "vel")}
        d["force"] = capi.DeviceArray((s.N, 4), np.float64)
        devs.append(d)
    refs = [coracle.cavity_force(s.pos, s.charge, s.image, s.box, s.L_typeid, 0.01, 1e-3) for s in systems]
    st = capi.Stream()
    for it in range(40):
        k = it % len(systems)
        s, d = systems[k], devs[k]
        a = capi.BussiArgs(KT, TAU, 0.0, 3.0 * (s.N - 1) - 3.0, 0.0, 1.0)  # dt = 0: velocities stay put
        handle.step(d["pos"], d["charge"], d["image"], d["force"], d["vel"], s.N, s.box, s.L_typeid, p, 0, s.N - 1, a, st.ptr)
        if it >= 36:
            f = d["force"].numpy(st.ptr)
            assert np.abs(f - refs[k]["force"]).max() <= 1e-10 * np.abs(refs[k]["force"]).max()


@pytest.mark.parametrize("n_mol", [1_000_000, 16_000_000])
def test_step_full_sizes(handle, coracle, n_mol):
    """BASELINE.json sizes (configs[1]: 1M, configs[3]: 16M), default tuning (folder step kernel on a full grid):
    oracle parity, plus the size-independent properties of the thermostat half -- the kinetic energy after the
    rescale is alpha^2 times the one before (to rounding), velocities keep their direction, masses untouched."""
    handle.set_tuning(variant=3, threads=384, ctas_per_sm=2, unroll=2)
    s = synth.make_system(n_mol)
    a = args_for(n_mol)
    out = run_step(handle, s, 0, n_mol, a)
    check_against_oracle(coracle, s, 0, n_mol, a, out)
    f, v, en, dip, ph, bo = out
    assert ph == n_mol and bo["err"] == 0.0
    ke0 = 0.5 * np.sum(s.vel[:n_mol, 3] * np.sum(s.vel[:n_mol, :3] ** 2, axis=1))
    ke1 = 0.5 * np.sum(v[:n_mol, 3] * np.sum(v[:n_mol, :3] ** 2, axis=1))
    assert abs(bo["ke"] - ke0) <= 1e-12 * ke0
    assert abs(ke1 - bo["alpha"] ** 2 * ke0) <= 1e-12 * ke0
    assert np.array_equal(v[:, 3], s.vel[:, 3])
    assert np.array_equal(v[n_mol], s.vel[n_mol])  # the photon is outside the thermostatted group


@pytest.mark.parametrize("n_mol", [262145, 500000, 1_000_000])
def test_step_automatic_cta_size(handle, coracle, n_mol):
    """Default tuning: the step kernel picks 320 / 352 / 384 threads per CTA from the particle count
    (fewest dependent memory round trips per pass); an explicit `threads` switches that off.  Parity either way."""
    s = synth.make_system(n_mol, replica=2)
    a = args_for(n_mol)
    h2 = capi.Handle(0)   # fresh handle: defaults
    try:
        assert h2.get_tuning("auto_threads") == 1 and h2.get_tuning("variant") == 3
        check_against_oracle(coracle, s, 0, n_mol, a, run_step(h2, s, 0, n_mol, a))
        h2.set_tuning(threads=384)
        assert h2.get_tuning("auto_threads") == 0
        check_against_oracle(coracle, s, 0, n_mol, a, run_step(h2, s, 0, n_mol, a))
    finally:
        h2.close()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("cooperative", [False, True])
def test_two_persistent_kernels_on_two_streams(coracle, cooperative):
    """Two handles on two unordered streams: the persistent grids of the two calls can hold each other's SM slots.
    Contract (include/cavb200.h, conventions):
      * cooperative launches (tuning pdl = 0): the driver co-schedules each grid -- every call is correct, no fault;
      * default (programmatic dependent launch): a call whose grid was not co-resident gives up after 50 ms, never
        hangs, never reports a wrong result as good (its *_read returns cudaErrorLaunchTimeout); the NEXT compute call
        on that handle returns cudaErrorLaunchTimeout once and the handle launches cooperatively from then on, so
        everything after the reported fault is correct."""
    n_mol = 400000
    s = synth.make_system(n_mol)
    ref = coracle.cavity_force(s.pos, s.charge, s.image, s.box, s.L_typeid, 0.01, 1e-3)
    p = capi.Params.make(0.01, 1e-3)
    hs = [capi.Handle(0), capi.Handle(0)]
    sts = [capi.Stream(), capi.Stream()]
    if cooperative:
        for h in hs:
            h.set_tuning(pdl=0)
    devs = []
    for _ in range(2):
        d = {k: capi.DeviceArray.from_numpy(getattr(s, k)) for k in ("pos", "charge", "image")}
        d["force"] = capi.DeviceArray.from_numpy(np.full((s.N, 4), np.nan))
        devs.append(d)
    flagged_reads = reported = 0
    try:
        for it in range(12):
            for k in range(2):
                d = devs[k]
                try:
                    hs[k].force(d["pos"], d["charge"], d["image"], d["force"], s.N, s.box, s.L_typeid, p, sts[k].ptr)
                except capi.CavbError as e:
                    # an EARLIER call of this handle timed out: reported once, handle now cooperative, this call redone
                    assert e.code == 702 and not cooperative
                    reported += 1
                    assert hs[k].get_tuning("pdl") == 0 and hs[k].fault_count >= 1
                    hs[k].force(d["pos"], d["charge"], d["image"], d["force"], s.N, s.box, s.L_typeid, p, sts[k].ptr)
            for k in range(2):
                try:
                    en, dip, ph = hs[k].force_read(sts[k].ptr)
                except capi.CavbError as e:
                    assert e.code == 702 and not cooperative and hs[k].get_tuning("pdl") == 1
                    flagged_reads += 1  # this very call timed out: reported here too, not silently wrong
                    continue
                f = devs[k]["force"].numpy(sts[k].ptr)
                assert ph == ref["photon_idx"] and np.allclose(en, ref["energies"], rtol=1e-10)
                assert np.abs(f - ref["force"]).max() <= 1e-10 * np.abs(ref["force"]).max()
        if cooperative:
            assert flagged_reads == 0 and reported == 0 and all(h.fault_count == 0 for h in hs)
        print(f"two unordered streams (cooperative={cooperative}): {flagged_reads} of 24 reads flagged a hand-off timeout, "
              f"{reported} reported at the next call, none hung, none wrong")
    finally:
        capi.sync()
        for h in hs:
            h.close()


def test_error_flags_do_not_stick(handle, coracle):
    """The status words are rewritten by every call: after a zero-kinetic-energy event (err = 1, the reference throws,
    src/BussiReservoirThermostat.h:57-61) the next healthy call reads err = 0 and the cumulative reservoir energy is
    untouched by the failed call."""
    n = 5000
    s = synth.make_system(n)
    a = capi.BussiArgs(KT, TAU, DT, 3.0 * n - 3.0, 0.2, (3.0 * n - 4.0) / 2)
    d_v = capi.DeviceArray.from_numpy(s.vel)
    handle.bussi_reset()
    handle.bussi(d_v, None, 0, n, a)
    good = handle.bussi_read()
    assert good["err"] == 0.0 and good["cumulative"] != 0.0
    d_zero = capi.DeviceArray.from_numpy(np.concatenate([np.zeros((s.N, 3)), s.vel[:, 3:]], axis=1))
    handle.bussi(d_zero, None, 0, n, a)
    bad = handle.bussi_read()
    assert bad["err"] == 1.0 and bad["cumulative"] == good["cumulative"]
    handle.bussi(d_v, None, 0, n, a)
    again = handle.bussi_read()
    assert again["err"] == 0.0 and again["cumulative"] != good["cumulative"]
    handle.bussi_reset()


def test_step_randomised_cases(coracle):
    """40 seeded random cases through the default tuning (fresh handle): sizes from 1 to 400k (both sides of the
    full-grid boundary where the folder kernel takes over), photon placement, thermostat window, charge pattern,
    coupling and draws all random; each against the oracle."""
    h = capi.Handle(0)
    rng = np.random.default_rng(20261018)
    try:
        for case in range(40):
            n_mol = int(rng.choice([rng.integers(1, 400), rng.integers(400, 20000), rng.integers(20000, 400000)]))
            photon = str(rng.choice(["last", "first", "middle", "absent", "duplicated"]))
            charges = str(rng.choice(["neutral", "nonneutral", "zero"]))
            s = synth.make_system(n_mol, replica=case, photon=photon, charges=charges)
            if rng.random() < 0.5:
                first, n = 0, (n_mol if photon in ("last", "absent") else s.N)
            else:
                first = int(rng.integers(0, s.N))
                n = int(rng.integers(0, s.N - first + 1))
            g = float(10.0 ** rng.uniform(-4, -2))
            omegac = float(rng.uniform(0.005, 0.02))
            dof = max(3.0 * n - 3.0, 0.0)
            a = capi.BussiArgs(KT, float(rng.choice([TAU, 100.0, 0.0])), DT, dof, float(rng.normal()),
                               float(rng.gamma(max(dof - 1.0, 2.0) / 2.0)))
            out = run_step(h, s, first, n, a, omegac=omegac, g=g)
            check_against_oracle(coracle, s, first, n, a, out, omegac=omegac, g=g)
            assert out[5]["err"] == 0.0 or n == 0, (case, n_mol, photon, first, n)
    finally:
        h.close()


@pytest.mark.parametrize("photon", ["last", "first", "middle", "absent", "duplicated"])
@pytest.mark.parametrize("n_mol", [1, 2, 33, 385, 500, 799])
def test_single_cta_kernel_for_small_systems(coracle, n_mol, photon):
    """Calls over at most `small_n` particles (default 768; the reference's own example has 501) run as ONE CTA without
    any inter-CTA hand-off (k_small).  Same results as the multi-CTA persistent kernels (tuning small_n = 0) -- the dipole
    to the last bit or two (another partition of the same compensated sum), everything else 1e-14 -- both against the
    oracle; the one-launch step stays bit-identical to the two calls; index-list and windowed groups included."""
    h = capi.Handle(0)
    try:
        s = synth.make_system(n_mol, replica=n_mol, photon=photon)
        first, n = (0, n_mol) if photon in ("last", "absent") else (0, s.N)
        a = args_for(n)
        outs = {}
        h.set_tuning(cluster_n=0)  # (small_n = 0 alone would send these sizes to the single-cluster kernel, tested below)
        for small in (0, 800):
            h.set_tuning(small_n=small)
            outs[small] = run_step(h, s, first, n, a)
            check_against_oracle(coracle, s, first, n, a, outs[small])
        f0, v0, en0, dip0, ph0, bo0 = outs[0]
        f1, v1, en1, dip1, ph1, bo1 = outs[800]
        assert ph0 == ph1 and np.abs(f1 - f0).max() <= 1e-14 * max(np.abs(f0).max(), 1e-300)
        assert np.allclose(v1, v0, rtol=1e-14, atol=0) and abs(bo1["alpha"] - bo0["alpha"]) <= 1e-14
        # the step == the two calls, bit for bit, on the single-CTA path too
        h.set_tuning(small_n=800)
        p = capi.Params.make(0.01, 1e-3)
        dev = {k: capi.DeviceArray.from_numpy(getattr(s, k)) for k in ("pos", "charge", "image", "vel")}
        d_f = capi.DeviceArray.from_numpy(np.full((s.N, 4), np.nan))
        h.bussi_reset()
        h.force(dev["pos"], dev["charge"], dev["image"], d_f, s.N, s.box, s.L_typeid, p)
        h.bussi(dev["vel"], None, first, n, a)
        assert np.array_equal(d_f.numpy().view(np.uint64), f1.view(np.uint64))
        assert np.array_equal(dev["vel"].numpy().view(np.uint64), v1.view(np.uint64))
        assert h.bussi_read()["alpha"] == bo1["alpha"] and np.array_equal(h.force_read()[0], en1)
        # the group as an index list (the molecular particles), and the kinetic energy alone
        idx = synth.molecular_group(s)
        if len(idx):
            d_idx = capi.DeviceArray.from_numpy(idx)
            d_v = capi.DeviceArray.from_numpy(s.vel)
            al = args_for(len(idx))
            vref = s.vel.copy()
            alpha_ref, ke_ref = coracle.bussi_step(vref, idx, al.dof, DT, KT, TAU, al.r_normal, al.gamma_draw, np.zeros(2))
            h.bussi_ke(d_v, d_idx, 0, len(idx))
            assert abs(h.bussi_read()["ke"] - ke_ref) <= 1e-12 * ke_ref
            h.bussi(d_v, d_idx, 0, len(idx), al)
            assert abs(h.bussi_read()["alpha"] - alpha_ref) <= 1e-12 * abs(alpha_ref)
            assert np.allclose(d_v.numpy(), vref, rtol=1e-12, atol=0)
    finally:
        h.close()


@pytest.mark.parametrize("photon", ["last", "first", "middle", "absent", "duplicated"])
@pytest.mark.parametrize("n_mol,ctas", [(1025, 16), (1537, 8), (3000, 16), (4096, 8), (8191, 16)])
def test_single_cluster_kernel_for_mid_size_systems(coracle, n_mol, ctas, photon):
    """Calls that include the force, over more than `small_n` and at most `cluster_n` particles (default 8192), run as ONE
    thread-block cluster (k_cluster): the CTAs' records meet behind the hardware cluster barrier instead of being polled.
    Same fold as the persistent kernels, so the same results as with tuning cluster_n = 0 (1e-14; bitwise wherever the
    partition into CTAs happens to coincide), both checked against the oracle; the one-launch step equals the force call
    followed by the Bussi call bit for bit in the force, and the rank-1 call leaves the same scalars."""
    h = capi.Handle(0)
    try:
        s = synth.make_system(n_mol, replica=n_mol, photon=photon)
        first, n = (0, n_mol) if photon in ("last", "absent") else (0, s.N)
        a = args_for(n)
        outs = {}
        for cl in (0, 8192):
            h.set_tuning(cluster_n=cl, cluster_ctas=ctas)
            outs[cl] = run_step(h, s, first, n, a)
            check_against_oracle(coracle, s, first, n, a, outs[cl])
        f0, v0, en0, dip0, ph0, bo0 = outs[0]
        f1, v1, en1, dip1, ph1, bo1 = outs[8192]
        assert ph0 == ph1 and np.abs(f1 - f0).max() <= 1e-14 * max(np.abs(f0).max(), 1e-300)
        assert np.allclose(v1, v0, rtol=1e-14, atol=0) and abs(bo1["alpha"] - bo0["alpha"]) <= 1e-14
        assert np.allclose(en1, en0, rtol=1e-13, atol=0) and np.allclose(dip1, dip0, rtol=1e-13, atol=1e-300)
        # the force call alone takes the cluster kernel too; its forces are those of the one-launch step
        p = capi.Params.make(0.01, 1e-3)
        dev = {k: capi.DeviceArray.from_numpy(getattr(s, k)) for k in ("pos", "charge", "image", "vel")}
        d_f = capi.DeviceArray.from_numpy(np.full((s.N, 4), np.nan))
        launches = h.launch_count
        h.force(dev["pos"], dev["charge"], dev["image"], d_f, s.N, s.box, s.L_typeid, p)
        assert h.launch_count == launches + 1
        assert np.array_equal(d_f.numpy().view(np.uint64), f1.view(np.uint64))
        assert np.array_equal(h.force_read()[0], en1)
        # rank-1 mode (no force array): same scalars
        h.force_rank1(dev["pos"], dev["charge"], dev["image"], s.N, s.box, s.L_typeid, p)
        en_r, dip_r, ph_r = h.force_read()
        assert ph_r == ph1 and np.array_equal(en_r, en1) and np.array_equal(dip_r, dip1)
        assert h.fault_count == 0
    finally:
        h.close()


def test_cluster_kernel_back_to_back_launches_of_different_shapes():
    """600 force calls back to back on one stream, alternating between three systems whose calls take a cluster of 8
    (1500 particles), a cluster of 16 (6000) and the single CTA (400): consecutive launches overlap under programmatic
    dependent launch, each with its own shared-memory inboxes.  Every launch must reproduce the bits of its system's
    first result."""
    h = capi.Handle(0)
    try:
        p = capi.Params.make(0.01, 1e-3)
        st = capi.Stream()
        systems = []
        for n_mol in (1500, 6000, 400):
            s = synth.make_system(n_mol, replica=n_mol)
            d = {k: capi.DeviceArray.from_numpy(getattr(s, k)) for k in ("pos", "charge", "image")}
            outs = [capi.DeviceArray.from_numpy(np.full((s.N, 4), np.nan)) for _ in range(4)]
            h.force(d["pos"], d["charge"], d["image"], outs[0], s.N, s.box, s.L_typeid, p, st.ptr)
            first = (outs[0].numpy(st.ptr).copy(), h.force_read(st.ptr)[0].copy())
            systems.append((s, d, outs, first))
        for k in range(600):
            s, d, outs, first = systems[k % 3]
            h.force(d["pos"], d["charge"], d["image"], outs[1 + (k // 3) % 3], s.N, s.box, s.L_typeid, p, st.ptr)
        st.sync()
        for s, d, outs, first in systems:
            for o in outs[1:]:
                assert np.array_equal(o.numpy(st.ptr).view(np.uint64), first[0].view(np.uint64))
        # the scalars left behind are those of the last call (system index 599 % 3 = 2)
        assert np.array_equal(h.force_read(st.ptr)[0], systems[2][3][1])
        assert h.fault_count == 0
    finally:
        h.close()


@pytest.mark.parametrize("n_mol", [40000, 131072, 350000, 500000, 1200000])
def test_launch_shape_rule_changes_nothing_but_the_partition(coracle, n_mol):
    """Below ~2M particles the launcher picks one CTA per SM for some calls (hotpath.cu launch_u: force call 704 / 768
    threads, Bussi call and step 512); an explicit `threads` switches the rule off.  Force call, Bussi call and the
    one-launch step under the rule and under two 384-thread CTAs per SM: same photon, forces / energies / alpha to
    1e-14 of each other (another partition of the same compensated sums), and the rule's results against the oracle."""
    s = synth.make_system(n_mol, replica=7)
    a = args_for(n_mol)
    p = capi.Params.make(0.01, 1e-3)
    res = {}
    for rule in (True, False):
        h = capi.Handle(0)
        try:
            if not rule:
                h.set_tuning(threads=384)
            step = run_step(h, s, 0, n_mol, a)
            dev = {k: capi.DeviceArray.from_numpy(getattr(s, k)) for k in ("pos", "charge", "image", "vel")}
            d_f = capi.DeviceArray.from_numpy(np.full((s.N, 4), np.nan))
            h.bussi_reset()
            h.force(dev["pos"], dev["charge"], dev["image"], d_f, s.N, s.box, s.L_typeid, p)
            h.bussi(dev["vel"], None, 0, n_mol, a)
            res[rule] = (step, d_f.numpy(), dev["vel"].numpy(), h.force_read(), h.bussi_read())
            assert h.fault_count == 0
        finally:
            h.close()
    check_against_oracle(coracle, s, 0, n_mol, a, res[True][0])
    (st1, f1, v1, fr1, b1), (st0, f0, v0, fr0, b0) = res[True], res[False]
    scale = np.abs(f0).max()
    assert fr1[2] == fr0[2] == st1[4] == st0[4]
    assert np.abs(f1 - f0).max() <= 1e-14 * scale and np.abs(st1[0] - st0[0]).max() <= 1e-14 * scale
    assert np.allclose(fr1[0], fr0[0], rtol=1e-13, atol=0) and np.allclose(st1[2], st0[2], rtol=1e-13, atol=0)
    assert abs(b1["alpha"] - b0["alpha"]) <= 1e-14 and abs(st1[5]["alpha"] - st0[5]["alpha"]) <= 1e-14
    assert np.allclose(v1, v0, rtol=1e-14, atol=0) and np.allclose(st1[1], st0[1], rtol=1e-14, atol=0)
    # the two calls against the one launch, under the rule: force call and step take different shapes in places
    assert np.abs(f1 - st1[0]).max() <= 1e-14 * scale and np.allclose(v1, st1[1], rtol=1e-14, atol=0)
