"""Sharded step on ONE GPU (nranks = 1): the reduce / exchange / apply kernels of shard.cu must give
exactly what the single-GPU step gives, including the global photon index with an index offset.
The multi-rank exchange itself is covered by tools/shard_check.py (run with torchrun on >= 2 GPUs)
and by the world_size-2 gloo test of the host-side combine order (tests/test_shard_host.py)."""
import numpy as np
import pytest

from cav_hoomd_b200 import capi, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_mol", [1000, 262145])
def test_shard_step_single_rank_equals_step(coracle, n_mol):
    h = capi.Handle(0)
    s = synth.make_system(n_mol)
    p = capi.Params.make(0.01, 1e-3)
    dof = 3.0 * n_mol - 3
    a = capi.BussiArgs(synth.KT_100K, synth.TAU_5PS, synth.DT_1FS, dof, -0.4, (dof - 1) / 2)
    dev = {k: capi.DeviceArray.from_numpy(getattr(s, k)) for k in ("pos", "charge", "image", "vel")}
    d_f = capi.DeviceArray.from_numpy(np.full((s.N, 4), np.nan))
    for _ in range(3):  # several steps: the mailbox parity / sequence logic is exercised
        dev["vel"].upload(s.vel)
        h.bussi_reset()
        h.shard_step(dev["pos"], dev["charge"], dev["image"], d_f, dev["vel"], s.N, 0, s.box, s.L_typeid, p, 0, n_mol, a)
        en, dip, ph = h.force_read()
        bo = h.bussi_read()
    ref = coracle.cavity_force(s.pos, s.charge, s.image, s.box, s.L_typeid, 0.01, 1e-3)
    assert ph == ref["photon_idx"]
    assert np.abs(d_f.numpy() - ref["force"]).max() <= 1e-10 * np.abs(ref["force"]).max()
    assert np.allclose(en, ref["energies"], rtol=1e-10)
    vref = s.vel.copy()
    alpha, ke = coracle.bussi_step(vref, np.arange(n_mol, dtype=np.uint32), dof, synth.DT_1FS, synth.KT_100K,
                                   synth.TAU_5PS, a.r_normal, a.gamma_draw, np.zeros(2))
    assert abs(bo["alpha"] - alpha) <= 1e-12 * abs(alpha)
    assert np.allclose(dev["vel"].numpy(), vref, rtol=1e-12, atol=0)
    h.close()


def test_shard_index_offset(coracle):
    """A shard that starts at global index 5000: the reported photon index is global, the force of
    the local photon row is the photon force."""
    h = capi.Handle(0)
    s = synth.make_system(2000, photon="middle")
    p = capi.Params.make(0.01, 1e-3)
    a = capi.BussiArgs(synth.KT_100K, synth.TAU_5PS, 0.0, 1.0, 0.0, 0.0)  # deltaT = 0: no rescale
    dev = {k: capi.DeviceArray.from_numpy(getattr(s, k)) for k in ("pos", "charge", "image", "vel")}
    d_f = capi.DeviceArray((s.N, 4), np.float64)
    h.shard_step(dev["pos"], dev["charge"], dev["image"], d_f, dev["vel"], s.N, 5000, s.box, s.L_typeid, p, 0, 0, a)
    en, dip, ph = h.force_read()
    ref = coracle.cavity_force(s.pos, s.charge, s.image, s.box, s.L_typeid, 0.01, 1e-3)
    assert ph == 5000 + ref["photon_idx"]
    assert np.abs(d_f.numpy() - ref["force"]).max() <= 1e-10 * np.abs(ref["force"]).max()
    assert np.array_equal(dev["vel"].numpy(), s.vel)
    h.close()


def test_shard_step_launch_modes_bitwise():
    """The one-kernel sharded step under programmatic dependent launch (default) and under the cooperative attribute
    (tuning pdl=0, what a handle falls back to after a missed hand-off): same bits, several steps back to back on one
    stream so that consecutive launches really overlap their tails."""
    n_mol = 100_003
    s = synth.make_system(n_mol)
    p = capi.Params.make(0.01, 1e-3)
    dof = 3.0 * n_mol - 3
    a = capi.BussiArgs(synth.KT_100K, synth.TAU_5PS, synth.DT_1FS, dof, 0.3, (dof - 1) / 2)
    out = {}
    for pdl in (1, 0):
        h = capi.Handle(0)
        h.set_tuning(pdl=pdl)
        dev = {k: capi.DeviceArray.from_numpy(getattr(s, k)) for k in ("pos", "charge", "image", "vel")}
        d_f = capi.DeviceArray.from_numpy(np.full((s.N, 4), np.nan))
        h.bussi_reset()
        for _ in range(5):  # velocities are rescaled five times over: every launch reads what the previous one wrote
            h.shard_step(dev["pos"], dev["charge"], dev["image"], d_f, dev["vel"], s.N, 0, s.box, s.L_typeid, p, 0, n_mol, a)
        out[pdl] = (d_f.numpy(), dev["vel"].numpy(), h.force_read()[0], h.bussi_read())
        assert h.fault_count == 0
        h.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    assert np.array_equal(out[0][2], out[1][2]) and out[0][3] == out[1][3]
    assert not np.array_equal(out[1][1], s.vel)
