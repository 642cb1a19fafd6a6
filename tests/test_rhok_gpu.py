"""GPU parity of the batched F(k,t) density field against the NumPy restatement of
reference src/cavitymd/analysis.py:34-47 (compute_density_field) and :359-364.

Tolerance: each term cos/sin(k.r) carries ~|k.r| eps ~ 5e-14 of argument rounding (|k.r| up to
~500 for a 1M-particle box) and the two sums run in different orders (NumPy pairwise vs fixed
per-thread strides), so |d rho_k| <= 1e-13 * N absolute (1e-10 of the typical |rho_k| ~ sqrt(N)).
"""
import numpy as np
import pytest

from cav_hoomd_b200 import capi, synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def gpu_rhok(handle, frames, kvec, stride=3):
    T, N = frames.shape[0], frames.shape[1]
    if stride == 4:
        f4 = np.zeros((T, N, 4))
        f4[:, :, :3] = frames
        f4[:, :, 3] = 7.0  # type bits must be ignored
        frames = f4
    d_pos = capi.DeviceArray.from_numpy(frames)
    d_k = capi.DeviceArray.from_numpy(kvec)
    d_rho = capi.DeviceArray((T, len(kvec), 2), np.float64)
    d_rho.fill_bytes(0xFF)
    handle.rhok(d_pos, stride, N * stride, N, T, d_k, len(kvec), d_rho)
    r = d_rho.numpy()
    return r[..., 0] + 1j * r[..., 1], d_rho


@pytest.mark.parametrize("stride", [3, 4])
@pytest.mark.parametrize("n_mol,K,T", [(1, 1, 1), (33, 50, 3), (1000, 64, 5), (5000, 300, 2), (100001, 64, 4)])
def test_rhok_matches_numpy(handle, n_mol, K, T, stride):
    s = synth.make_system(n_mol)
    frames = synth.random_walk_frames(s, T)
    kvec = synth.fibonacci_sphere(max(K, 2))[:K] * 1.0
    rho, _ = gpu_rhok(handle, frames, kvec, stride)
    for t in range(T):
        ref = O.numpy_density_field(frames[t], kvec)
        assert np.abs(rho[t] - ref).max() <= 1e-13 * s.N + 1e-12


def test_rhok_long_double_truth_and_fkt(handle, coracle):
    s = synth.make_system(20000)
    T, K = 12, 64
    frames = synth.random_walk_frames(s, T, sigma=0.5)
    kvec = synth.fibonacci_sphere(K)
    rho, d_rho = gpu_rhok(handle, frames, kvec)
    truth = coracle.rhok(frames[3], kvec)
    assert np.abs(rho[3] - truth).max() <= 1e-13 * s.N
    n_or, n_lag = 8, 6
    d_out = capi.DeviceArray((n_or, n_lag), np.float64)
    handle.fkt(d_rho, T, K, n_or, n_lag, d_out)
    F = d_out.numpy()
    for o in range(n_or):
        for l in range(n_lag):
            if o + l < T:
                ref = O.numpy_field_autocorr(rho[o], rho[o + l])
                assert abs(F[o, l] - ref) <= 1e-12 * max(abs(ref), 1.0)
            else:
                assert np.isnan(F[o, l])
    # F(k, 0) = mean_k |rho_k|^2 >= 0 (size-independent property)
    assert np.all(F[:, 0] > 0)


def test_rhok_deterministic(handle):
    s = synth.make_system(50000)
    frames = synth.random_walk_frames(s, 2)
    kvec = synth.fibonacci_sphere(64)
    a, _ = gpu_rhok(handle, frames, kvec)
    b, _ = gpu_rhok(handle, frames, kvec)
    assert np.array_equal(a, b)


def test_fkt_from_a_gsd_trajectory(tmp_path, handle):
    """BASELINE configs[4] is phrased on a GSD trajectory (--enable-fkt): frames written and read back with the
    minimal GSD layer (cav_hoomd_b200/gsdio.py; float32 positions, as HOOMD stores them) feed the batched
    density-field kernel; compared with the NumPy restatement of compute_density_field on the same frames
    (reference src/cavitymd/analysis.py:34-47).  Tolerance 1e-9 relative to N (device sincos vs libm)."""
    from cav_hoomd_b200 import analysis, gsdio, synth
    from oracle import oracle as O
    s = synth.make_system(3000, replica=5)
    rng = np.random.default_rng(0)
    frames = []
    for t in range(6):
        fr = gsdio.system_to_frame(s, step=100 * t)
        fr.position = (fr.position + rng.normal(scale=0.05 * t, size=fr.position.shape)).astype(np.float32)
        frames.append(fr)
    path = str(tmp_path / "traj.gsd")
    gsdio.write_gsd(path, frames)
    with gsdio.open_gsd(path) as f:
        assert len(f) == 6 and [f[i].step for i in range(6)] == [0, 100, 200, 300, 400, 500]
        pos = np.stack([f[i].position for i in range(len(f))])   # float32, as stored: DensityField takes the f32 path
        assert pos.dtype == np.float32
    kvec = analysis.generate_fibonacci_sphere(16) * 0.7
    field = analysis.DensityField(kvec, handle)
    rho = field.compute_density_field(pos)
    for t in range(6):
        ref = O.numpy_density_field(pos[t], kvec)
        assert np.abs(rho[t] - ref).max() <= 1e-9 * pos.shape[1]
    f0 = O.numpy_field_autocorr(rho[0], rho[3])
    d_rho = capi.DeviceArray.from_numpy(np.stack([rho.real, rho.imag], axis=-1))
    F = field.autocorrelation(d_rho, 6, 2, 4)
    assert abs(F[0, 3] - f0) <= 1e-9 * abs(f0)


@pytest.mark.parametrize("N,T,K", [(1, 1, 1), (1000, 3, 7), (5001, 2, 64), (150000, 4, 33)])
def test_rhok_float32_positions(handle, N, T, K):
    """cavb200_rhok_f32 (GSD-style float32 xyz): bit-identical to cavb200_rhok on the same values widened on the host
    (the kernel widens exactly), and within 1e-9*N of the NumPy restatement np.dot(float32 positions, float64 k)."""
    rng = np.random.default_rng(N + T)
    pos32 = (rng.uniform(-40, 40, size=(T, N, 3))).astype(np.float32)
    kvec = O.numpy_fibonacci_sphere(max(K, 2))[:K] * 0.9
    d_k = capi.DeviceArray.from_numpy(kvec)
    d32 = capi.DeviceArray.from_numpy(pos32)
    d64 = capi.DeviceArray.from_numpy(pos32.astype(np.float64))
    r32 = capi.DeviceArray((T, K, 2), np.float64)
    r64 = capi.DeviceArray((T, K, 2), np.float64)
    handle.rhok_f32(d32, 3 * N, N, T, d_k, K, r32)
    handle.rhok(d64, 3, 3 * N, N, T, d_k, K, r64)
    a, b = r32.numpy(), r64.numpy()
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64))
    for t in range(T):
        ref = O.numpy_density_field(pos32[t], kvec)
        assert np.abs((a[t, :, 0] + 1j * a[t, :, 1]) - ref).max() <= 1e-9 * N
    with pytest.raises(capi.CavbError):
        handle.rhok(d64, 13, 3 * N, N, T, d_k, K, r64)  # the internal float32 code is not a public stride


@pytest.mark.parametrize("kmag", [3.0e4, 1.0e7, 1.0e12])
def test_rhok_huge_arguments_take_the_library_path(handle, kmag):
    """|k.r| >= 2^20 (and any inf / nan) leaves the inline two-constant Cody-Waite reduction for the library's
    Payne-Hanek sincos (csrc/rhok.cu, `huge` branch) -- per particle, so a frame may mix both.  Against NumPy (libm does
    an exact reduction as well): each term is exact to ~1 ulp of the ARGUMENT's rounding, |k.r| eps, on both sides, so
    the sums agree to N |k.r|max eps."""
    n_mol, K, T = 4000, 12, 2
    s = synth.make_system(n_mol, replica=8)
    frames = synth.random_walk_frames(s, T)
    kvec = synth.fibonacci_sphere(K) * kmag
    kr = np.abs(frames[0] @ kvec.T)
    big = (kr >= 2.0 ** 20).mean()
    assert big > 0.5 if kmag >= 1e7 else 0.0 < big < 1.0  # 3e4: a genuine mix of both branches
    rho, _ = gpu_rhok(handle, frames, kvec, 3)
    for t in range(T):
        ref = O.numpy_density_field(frames[t], kvec)
        # identical arguments on both sides (same FMA-free dot? no: the device contracts k.r to FMAs, NumPy's BLAS too,
        # but not necessarily in the same order): allow the argument's own rounding, |k.r| * 2^-52, per term
        tol = s.N * np.abs(frames[t] @ kvec.T).max() * 2.0 ** -52 * 4 + 1e-12
        assert np.abs(rho[t] - ref).max() <= tol
        assert np.all(np.abs(rho[t]) <= s.N)


def test_rhok_nonfinite_positions_propagate(handle):
    """inf / nan coordinates go through the library branch too and come out as nan, as in NumPy (cos(inf) = nan)."""
    frames = np.zeros((1, 64, 3))
    frames[0, 5, 0] = np.inf
    kvec = synth.fibonacci_sphere(4) * 1.0
    rho, _ = gpu_rhok(handle, frames, kvec, 3)
    with np.errstate(invalid="ignore"):
        ref = O.numpy_density_field(frames[0], kvec)
    assert np.array_equal(np.isnan(rho[0].real), np.isnan(ref.real)) and np.isnan(rho[0].real).any()


def test_sincos_accuracy_per_argument(handle):
    """One particle per frame makes rho[t][k] = (cos, sin)(k . r_t): 64 x 1024 arguments spread over |k.r| < 2^19, each
    against long-double truth of the SAME rounded argument.  The kernel's own reduction, short polynomials and 512-entry
    table (rhok.cu sincos_reduced, the rotation by the table entry) must stay within 1 ulp of 1 -- absolute 2.3e-16 --
    for every one of them (worst case of the scheme by rounding analysis: 2.4e-16; NumPy simulation of the same
    operation order over 6e5 arguments: 2.16e-16)."""
    rng = np.random.default_rng(5)
    T, K = 1024, 64
    frames = np.zeros((T, 1, 3))
    frames[:, 0, 0] = rng.uniform(-1.0, 1.0, T) * 10.0 ** rng.uniform(-3, 5.7, T)
    kvec = np.zeros((K, 3))
    kvec[:, 0] = rng.uniform(0.5, 1.0, K)
    d_rho = capi.DeviceArray((T, K, 2), np.float64)
    handle.rhok(capi.DeviceArray.from_numpy(frames), 3, 3, 1, T, capi.DeviceArray.from_numpy(kvec), K, d_rho)
    rho = d_rho.numpy()
    arg = frames[:, 0, 0][:, None] * kvec[:, 0][None, :]  # one multiplication, rounded once, as in the kernel
    assert np.abs(arg).max() < 2.0 ** 19 and np.abs(arg).max() > 1e5
    la = arg.astype(np.longdouble)
    err_c = np.abs(rho[..., 0].astype(np.longdouble) - np.cos(la)).max()
    err_s = np.abs(rho[..., 1].astype(np.longdouble) - np.sin(la)).max()
    assert float(err_c) <= 2.3e-16 and float(err_s) <= 2.3e-16, (float(err_c), float(err_s))


def test_sincos_table_cell_edges(handle):
    """The table scheme's worst arguments: the edges of a table cell (|r| = pi/512, where the truncated polynomials are
    at their largest error and rint() may go either way), the cell centres (r = 0: the table entry itself) and the
    axes (multiples of pi/2, where one component is ~1e-16 and the table entry is an exact 0 or +-1).  Same bound as
    the random arguments: absolute 2.3e-16 against long-double truth of the rounded argument."""
    step = 2.0 * np.pi / 512.0
    cells = np.concatenate([np.arange(-1100, 1100), np.array([2 ** 17 + 5, -(2 ** 19) + 77, 40_000_001, -39_999_999])])
    offs = np.array([0.0, 0.5, 0.5 - 1e-9, 0.5 + 1e-9, -0.5, 0.25, 1e-12, -1e-12])
    args = ((cells[:, None] + offs[None, :]) * step).ravel()
    args = np.concatenate([args, np.arange(-64, 65) * (np.pi / 2.0)])
    assert np.abs(args).max() < 2.0 ** 19
    T = len(args)
    frames = np.zeros((T, 1, 3))
    frames[:, 0, 1] = args
    kvec = np.array([[0.0, 1.0, 0.0], [0.0, -1.0, 0.0], [0.0, 0.5, 0.0]])
    K = len(kvec)
    d_rho = capi.DeviceArray((T, K, 2), np.float64)
    handle.rhok(capi.DeviceArray.from_numpy(frames), 3, 3, 1, T, capi.DeviceArray.from_numpy(kvec), K, d_rho)
    rho = d_rho.numpy()
    la = (args[:, None] * kvec[:, 1][None, :]).astype(np.longdouble)  # exact products (powers of two)
    err_c = np.abs(rho[..., 0].astype(np.longdouble) - np.cos(la)).max()
    err_s = np.abs(rho[..., 1].astype(np.longdouble) - np.sin(la)).max()
    assert float(err_c) <= 2.3e-16 and float(err_s) <= 2.3e-16, (float(err_c), float(err_s))
    # exp(-ix) is the conjugate of exp(ix) bit for bit: the table is built by octant symmetry and rint() is odd
    assert np.array_equal(rho[:, 0, 0], rho[:, 1, 0]) and np.array_equal(rho[:, 0, 1], -rho[:, 1, 1])


@pytest.mark.parametrize("N", [511, 512, 513, 1537])
def test_rhok_many_frames_one_slice_per_frame(handle, N):
    """Thousands of frames in one launch make the grid rule cut a frame into ONE slice (rhok_slices: whole waves of
    resident CTAs), so a CTA walks all tiles of its frame through both tile buffers, ragged last tile included, and
    writes rho directly (no fold kernel) -- for the three position layouts, against the NumPy restatement."""
    T, K = 3000, 5
    rng = np.random.default_rng(N)
    pos32 = rng.uniform(-30, 30, size=(T, N, 3)).astype(np.float32)
    pos = pos32.astype(np.float64)
    kvec = O.numpy_fibonacci_sphere(8)[:K] * 1.1
    d_k = capi.DeviceArray.from_numpy(kvec)
    ref = np.exp(1j * np.einsum("tnc,kc->tkn", pos, kvec)).sum(axis=2)
    out = {}
    for name in ("xyz", "scalar4", "f32"):
        d_rho = capi.DeviceArray((T, K, 2), np.float64)
        d_rho.fill_bytes(0xFF)
        if name == "xyz":
            handle.rhok(capi.DeviceArray.from_numpy(pos), 3, 3 * N, N, T, d_k, K, d_rho)
        elif name == "scalar4":
            p4 = np.full((T, N, 4), np.nan)  # the fourth word (type bits) must never be looked at
            p4[:, :, :3] = pos
            handle.rhok(capi.DeviceArray.from_numpy(p4), 4, 4 * N, N, T, d_k, K, d_rho)
        else:
            handle.rhok_f32(capi.DeviceArray.from_numpy(pos32), 3 * N, N, T, d_k, K, d_rho)
        r = d_rho.numpy()
        out[name] = r
        assert np.abs((r[..., 0] + 1j * r[..., 1]) - ref).max() <= 1e-13 * N + 1e-12
    # same values, same walk order: the three layouts agree bit for bit
    assert np.array_equal(out["xyz"].view(np.uint64), out["scalar4"].view(np.uint64))
    assert np.array_equal(out["xyz"].view(np.uint64), out["f32"].view(np.uint64))
