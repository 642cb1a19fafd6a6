"""GPU results through the C ABI against the committed golden vectors (minted from the reference's
own translation units by tests/golden/make_golden.py).  Nothing here reads /root/reference."""
import os

import numpy as np
import pytest

from cav_hoomd_b200 import capi

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def cases(fname):
    g = np.load(os.path.join(GOLD, fname))
    # (the rot_* entries of bussi.npz are the host-side rotational factor: tests/test_plugin_gpu.py, tests/test_oracle.py)
    return sorted({k.split("/")[0] for k in g.files if "/" in k and not k.startswith("rot_")})


@pytest.mark.parametrize("case", cases("cavity_force.npz"))
def test_force_golden(handle, case):
    """forces / energies within 1e-10 relative of the reference's CPU class (BASELINE.json)."""
    g = np.load(os.path.join(GOLD, "cavity_force.npz"))
    omegac, gc, phmass, Lt = g[f"{case}/params"]
    pos, q, img = g[f"{case}/pos"], g[f"{case}/charge"], g[f"{case}/image"]
    N = pos.shape[0]
    d_f = capi.DeviceArray.from_numpy(np.full((N, 4), np.nan))
    handle.force(capi.DeviceArray.from_numpy(pos), capi.DeviceArray.from_numpy(q), capi.DeviceArray.from_numpy(img),
                 d_f, N, g[f"{case}/box"], int(Lt), capi.Params.make(omegac, gc, phmass))
    en, dip, ph = handle.force_read()
    f, fr = d_f.numpy(), g[f"{case}/force"]
    scale = np.abs(fr).max()
    if scale == 0:
        assert np.all(f == 0.0) and np.all(en == 0.0)
    else:
        assert np.abs(f - fr).max() <= 1e-10 * scale
        assert np.allclose(en, g[f"{case}/energies"], rtol=1e-10, atol=0)


@pytest.mark.parametrize("case", cases("bussi.npz"))
def test_bussi_golden(handle, case):
    """alpha, KE and reservoir bookkeeping over 4 consecutive steps against the reference's
    getRescalingFactorsOne (draws injected).  KE is summed in a different order: 1e-12 relative."""
    g = np.load(os.path.join(GOLD, "bussi.npz"))
    dof, kT, tau, dt = g[f"{case}/args"]
    vel0 = g[f"{case}/vel0"]
    n = int(round((dof + 3) / 3)) if dof > 1 else 64
    d_vel = capi.DeviceArray.from_numpy(vel0)
    handle.bussi_reset()
    for k, (r, gm) in enumerate(g[f"{case}/draws"]):
        handle.bussi(d_vel, None, 0, n, capi.BussiArgs(kT, tau, dt, dof, r, gm))
        out = handle.bussi_read()
        alpha, ke, cum, inst = g[f"{case}/rows"][k]
        assert abs(out["ke"] - ke) <= 1e-12 * ke
        assert abs(out["alpha"] - alpha) <= 1e-12 * abs(alpha)
        assert abs(out["instantaneous"] - inst) <= 1e-9 * max(abs(inst), 1e-12 * ke)
        assert abs(out["cumulative"] - cum) <= 1e-9 * max(abs(cum), 1e-12 * ke)
    assert np.allclose(d_vel.numpy(), g[f"{case}/vel_final"], rtol=1e-11, atol=0)


def test_fkt_golden(handle):
    g = np.load(os.path.join(GOLD, "fkt.npz"))
    frames, kvec = g["frames"], g["kvec"]
    T, N, _ = frames.shape
    d_rho = capi.DeviceArray((T, len(kvec), 2), np.float64)
    handle.rhok(capi.DeviceArray.from_numpy(frames), 3, 3 * N, N, T, capi.DeviceArray.from_numpy(kvec), len(kvec), d_rho)
    r = d_rho.numpy()
    assert np.abs((r[..., 0] + 1j * r[..., 1]) - g["rho"]).max() <= 1e-13 * N + 1e-12
    d_F = capi.DeviceArray((T, 4), np.float64)
    handle.fkt(d_rho, T, len(kvec), T, 4, d_F)
    F = d_F.numpy()
    ok = ~np.isnan(g["F"])
    assert np.array_equal(np.isnan(F), ~ok)
    assert np.allclose(F[ok], g["F"][ok], rtol=1e-11, atol=1e-9)
