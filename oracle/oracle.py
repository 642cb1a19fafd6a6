"""CPU oracle loader + NumPy restatements.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Nothing under cav_hoomd_b200/ does.

Three oracles, strongest first:
  RefOracle   oracle/_ref/libcavref.so -- the reference's own src/CavityForceCompute.cc and
              src/BussiReservoirThermostat.h compiled verbatim by path (oracle/ref_driver.cc).
  COracle     oracle/liboracle.so -- plain-C restatement (oracle/cavity_oracle.c).
  numpy_*     NumPy restatements of the reference's Python code paths
              (src/cavitymd/cavity_force_python.py:65-149, src/cavitymd/analysis.py:34-66,359-364).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_up = C.POINTER(C.c_uint32)


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _u(a):
    return None if a is None else a.ctypes.data_as(_up)


def build(force: bool = False) -> None:
    """make -C oracle (C restatement always; _ref only when /root/reference is present)."""
    if force or not os.path.exists(os.path.join(HERE, "liboracle.so")):
        subprocess.run(["make", "-C", HERE], check=True, stdout=subprocess.DEVNULL)


def have_ref(fma: bool = False) -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libcavref_fma.so" if fma else "libcavref.so"))


class COracle:
    """ctypes view of oracle/liboracle.so (cavity_oracle.h)."""

    def __init__(self, fma: bool = False):
        build()
        self.lib = C.CDLL(os.path.join(HERE, "liboracle_fma.so" if fma else "liboracle.so"))
        L = self.lib
        L.orc_find_photon.restype = C.c_int
        L.orc_find_photon.argtypes = [_dp, C.c_uint32, C.c_uint32]
        L.orc_cavity_force.restype = C.c_int
        L.orc_cavity_force.argtypes = [_dp, _dp, _ip, _dp, C.c_uint32, C.c_double, C.c_double, C.c_double,
                                       C.c_uint32, C.c_double, C.c_double, C.c_double, _dp, _dp]
        L.orc_dipole_exact.restype = None
        L.orc_dipole_exact.argtypes = [_dp, _dp, _ip, C.c_uint32, C.c_double, C.c_double, C.c_double,
                                       C.c_int, _dp]
        L.orc_bussi_rescale_factor.restype = C.c_double
        L.orc_bussi_rescale_factor.argtypes = [C.c_double] * 7
        L.orc_kinetic_energy.restype = C.c_double
        L.orc_kinetic_energy.argtypes = [_dp, _up, C.c_uint32]
        L.orc_rescale_velocities.restype = None
        L.orc_rescale_velocities.argtypes = [_dp, _up, C.c_uint32, C.c_double]
        L.orc_bussi_step.restype = C.c_double
        L.orc_bussi_step.argtypes = [_dp, _up, C.c_uint32, C.c_double, C.c_double, C.c_double, C.c_double,
                                     C.c_double, C.c_double, _dp, _dp]
        L.orc_rhok.restype = None
        L.orc_rhok.argtypes = [_dp, C.c_uint32, C.c_uint32, _dp, C.c_uint32, _dp, _dp]
        L.orc_nvt_step.restype = C.c_double
        L.orc_nvt_step.argtypes = [_dp, _dp, _dp, _ip, _dp, C.c_uint32, C.c_double, C.c_double, C.c_double, C.c_uint32,
                                   C.c_double, C.c_double, C.c_double, C.c_double, C.c_uint32, C.c_uint32, C.c_double,
                                   C.c_double, C.c_double, C.c_double, C.c_double, _dp, _dp, _dp]
        L.orc_nve_step.restype = None
        L.orc_nve_step.argtypes = [_dp, _dp, _dp, _ip, _dp, C.c_uint32, C.c_double, C.c_double, C.c_double,
                                   C.c_uint32, C.c_double, C.c_double, C.c_double, C.c_double, _dp]
        L.orc_nvt_step_w.restype = C.c_double
        L.orc_nvt_step_w.argtypes = L.orc_nvt_step.argtypes + [C.c_int]
        L.orc_nve_step_w.restype = None
        L.orc_nve_step_w.argtypes = L.orc_nve_step.argtypes + [C.c_int]
        L.orc_wrap.restype = None
        L.orc_wrap.argtypes = [_dp, _ip, _dp]

    def cavity_force(self, pos, charge, image, box, L_typeid, omegac, couplstr, phmass=1.0):
        """-> dict(force[N,4], energies[3], dipole[3], photon_idx)"""
        N = pos.shape[0]
        force = np.empty((N, 4))
        en = np.zeros(3)
        dip = np.zeros(3)
        ph = self.lib.orc_cavity_force(_d(pos), _d(charge), _i(image), _d(force), N, box[0], box[1], box[2],
                                       L_typeid, omegac, couplstr, phmass, _d(en), _d(dip))
        return dict(force=force, energies=en, dipole=dip, photon_idx=ph)

    def dipole_exact(self, pos, charge, image, box, photon_idx):
        dip = np.zeros(3)
        self.lib.orc_dipole_exact(_d(pos), _d(charge), _i(image), pos.shape[0], box[0], box[1], box[2],
                                  int(photon_idx), _d(dip))
        return dip

    def bussi_rescale_factor(self, K, dof, dt, kT, tau, r_normal, gamma_draw):
        return self.lib.orc_bussi_rescale_factor(K, dof, dt, kT, tau, r_normal, gamma_draw)

    def kinetic_energy(self, vel, idx=None):
        n = vel.shape[0] if idx is None else len(idx)
        return self.lib.orc_kinetic_energy(_d(vel), _u(idx), n)

    def bussi_step(self, vel, idx, dof, dt, kT, tau, r_normal, gamma_draw, reservoir):
        """In place on vel; reservoir = float64[2] {cumulative, instantaneous}. -> (alpha, KE)"""
        n = vel.shape[0] if idx is None else len(idx)
        ke = C.c_double(0.0)
        a = self.lib.orc_bussi_step(_d(vel), _u(idx), n, dof, dt, kT, tau, r_normal, gamma_draw,
                                    _d(reservoir), C.byref(ke))
        return a, ke.value

    def rhok(self, pos, kvec):
        pos = np.ascontiguousarray(pos)
        kvec = np.ascontiguousarray(kvec, dtype=np.float64)
        K = kvec.shape[0]
        re = np.zeros(K)
        im = np.zeros(K)
        self.lib.orc_rhok(_d(pos), pos.shape[1], pos.shape[0], _d(kvec), K, _d(re), _d(im))
        return re + 1j * im

    def nvt_step(self, pos, vel, charge, image, force, box, L_typeid, omegac, couplstr, phmass, dt, first, n, dof, kT,
                 tau, r_normal, gamma_draw, reservoir, ke_io, wrap=False):
        """One thermostatted harness step (orc_nvt_step_w); reservoir = float64[2], ke_io = float64[1], both
        updated in place; wrap: box wrap + image update after the drift (image is then updated in place).
        Returns (alpha, energies)."""
        en = np.zeros(3)
        alpha = self.lib.orc_nvt_step_w(_d(pos), _d(vel), _d(charge), _i(image), _d(force), pos.shape[0], box[0], box[1],
                                        box[2], L_typeid, omegac, couplstr, phmass, dt, first, n, dof, kT, tau, r_normal,
                                        gamma_draw, _d(reservoir), _d(ke_io), _d(en), int(wrap))
        return alpha, en

    def nve_step(self, pos, vel, charge, image, force, box, L_typeid, omegac, couplstr, phmass, dt, wrap=False):
        en = np.zeros(3)
        self.lib.orc_nve_step_w(_d(pos), _d(vel), _d(charge), _i(image), _d(force), pos.shape[0], box[0], box[1],
                                box[2], L_typeid, omegac, couplstr, phmass, dt, _d(en), int(wrap))
        return en

    def wrap(self, pos, image, box):
        """HOOMD BoxDim::wrap restated (orthorhombic, periodic): pos[N,>=3] and image[N,3] in place."""
        L = np.ascontiguousarray(box, dtype=np.float64)
        stride = pos.shape[1]
        for i in range(pos.shape[0]):
            p = np.ascontiguousarray(pos[i, :3])
            im = np.ascontiguousarray(image[i])
            self.lib.orc_wrap(_d(p), _i(im), _d(L))
            pos[i, :3] = p
            image[i] = im


class RefOracle:
    """ctypes view of oracle/_ref/libcavref.so: the reference's own code behind ref_driver.cc."""

    def __init__(self, fma: bool = False):
        path = os.path.join(HERE, "_ref", "libcavref_fma.so" if fma else "libcavref.so")
        if not os.path.exists(path):
            build(force=True)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        # the reference TU carries pybind11 export code -> needs libpython symbols: load globally
        self.lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
        L = self.lib
        L.ref_cavity_create.restype = C.c_void_p
        L.ref_cavity_create.argtypes = [C.c_uint32, C.c_double, C.c_double, C.c_double, C.c_uint32, C.c_uint32,
                                        C.c_double, C.c_double, C.c_double]
        L.ref_cavity_destroy.argtypes = [C.c_void_p]
        L.ref_cavity_load.argtypes = [C.c_void_p, _dp, _dp, _ip]
        L.ref_cavity_compute.restype = C.c_int
        L.ref_cavity_compute.argtypes = [C.c_void_p, C.c_uint32]
        L.ref_cavity_read.argtypes = [C.c_void_p, _dp, _dp]
        L.ref_bussi_create.restype = C.c_void_p
        L.ref_bussi_create.argtypes = [C.c_uint32, _dp, _up, C.c_uint32, C.c_double, C.c_double, C.c_double]
        L.ref_bussi_destroy.argtypes = [C.c_void_p]
        L.ref_bussi_step.restype = C.c_int
        L.ref_bussi_step.argtypes = [C.c_void_p, C.c_uint64, C.c_double, C.c_double, C.c_double, _dp]
        if hasattr(L, "ref_bussi_step_rot"):
            L.ref_bussi_step_rot.restype = C.c_int
            L.ref_bussi_step_rot.argtypes = [C.c_void_p, C.c_uint64, C.c_double, _dp, C.c_double, C.c_double, _dp]
        L.ref_bussi_read.argtypes = [C.c_void_p, _dp]
        L.ref_bussi_reset.argtypes = [C.c_void_p]

    # -- cavity force -------------------------------------------------------------------------
    def cavity_open(self, N, box, L_typeid, omegac, couplstr, phmass=1.0, ntypes=3):
        return self.lib.ref_cavity_create(N, box[0], box[1], box[2], L_typeid, ntypes, omegac, couplstr, phmass)

    def cavity_force(self, pos, charge, image, box, L_typeid, omegac, couplstr, phmass=1.0, ntypes=3):
        N = pos.shape[0]
        h = self.cavity_open(N, box, L_typeid, omegac, couplstr, phmass, ntypes)
        try:
            self.lib.ref_cavity_load(h, _d(pos), _d(charge), _i(image))
            rc = self.lib.ref_cavity_compute(h, 1)
            if rc != 0:
                raise RuntimeError("reference threw (no 'L' type?)")
            force = np.empty((N, 4))
            en = np.zeros(3)
            self.lib.ref_cavity_read(h, _d(force), _d(en))
        finally:
            self.lib.ref_cavity_destroy(h)
        return dict(force=force, energies=en)

    # -- Bussi ------------------------------------------------------------------------------
    def bussi_open(self, vel, idx, dof, kT, tau):
        n = vel.shape[0] if idx is None else len(idx)
        return self.lib.ref_bussi_create(vel.shape[0], _d(vel), _u(idx), n, dof, kT, tau)

    def bussi_step(self, h, timestep, dt, r_normal, gamma_draw):
        out = np.zeros(4)
        rc = self.lib.ref_bussi_step(h, timestep, dt, r_normal, gamma_draw, _d(out))
        if rc != 0:
            raise RuntimeError("Bussi thermostat requires non-zero initial momenta.")
        return dict(alpha=out[0], ke=out[1], cumulative=out[2], instantaneous=out[3])

    def bussi_step_rot(self, h, timestep, dt, draws, rot_dof, rot_ke):
        """getRescalingFactorsOne with rotational degrees of freedom (reference
        src/BussiReservoirThermostat.h:53-55,77-95); draws = [normal_t, gamma_t, normal_r, gamma_r]."""
        out = np.zeros(7)
        d = np.ascontiguousarray(draws, dtype=np.float64)
        rc = self.lib.ref_bussi_step_rot(h, timestep, dt, _d(d), rot_dof, rot_ke, _d(out))
        if rc < 0:
            raise RuntimeError("Bussi thermostat requires non-zero initial momenta.")
        return dict(alpha=out[0], alpha_rot=out[1], ke=out[2], cumulative=out[3], instantaneous=out[4],
                    cumulative_rot=out[5], instantaneous_rot=out[6], draws_left=rc)

    def bussi_read(self, h, N):
        v = np.empty((N, 4))
        self.lib.ref_bussi_read(h, _d(v))
        return v

    def bussi_close(self, h):
        self.lib.ref_bussi_destroy(h)


# ---------------------------------------------------------------------------------------------
# NumPy restatements of the reference's PYTHON code paths
# ---------------------------------------------------------------------------------------------
def numpy_unwrap(position, image, box_lengths):
    """src/cavitymd/utils.py unwrap_positions: position + image * box_lengths."""
    return position + image * np.asarray(box_lengths)


def numpy_cavity_force(pos, charge, image, box, cavity_typeid, omegac, couplstr, phmass=1.0):
    """src/cavitymd/cavity_force_python.py:65-149 restated (the reference's Python fallback).
    Differences from the C++ class, kept on purpose: the photon is INCLUDED in the
    dipole sum (:97) and every particle except the photon gets the molecular force (:131-135)."""
    from cav_hoomd_b200.synth import w_to_typeid

    K = phmass * omegac ** 2
    typeid = w_to_typeid(pos[:, 3])
    cav = np.where(typeid == cavity_typeid)[0]
    N = pos.shape[0]
    force = np.zeros((N, 3))
    if len(cav) == 0:
        return dict(force=force, energies=np.zeros(3))
    ci = cav[0]
    u = numpy_unwrap(pos[:, :3], image, box)
    d = np.dot(charge, u)
    dxy = d.copy()
    dxy[2] = 0.0
    q = u[ci]
    qxy = q.copy()
    qxy[2] = 0.0
    e = np.array([0.5 * K * np.dot(q, q), couplstr * np.dot(qxy, dxy), 0.5 * (couplstr ** 2 / K) * np.dot(dxy, dxy)])
    ff = qxy + (couplstr / K) * dxy
    force[:] = (-couplstr * charge)[:, None] * ff[None, :]
    force[ci] = -K * q - couplstr * dxy
    return dict(force=force, energies=e, dipole=d)


def numpy_fibonacci_sphere(samples=100):
    """src/cavitymd/analysis.py:50-66, loop form as written there."""
    points = np.zeros((samples, 3))
    phi = np.pi * (3.0 - np.sqrt(5.0))
    for i in range(samples):
        y = 1 - (i / float(samples - 1)) * 2
        radius = np.sqrt(1 - y * y)
        theta = phi * i
        points[i] = [np.cos(theta) * radius, y, np.sin(theta) * radius]
    return points


def numpy_density_field(positions, wavevectors):
    """src/cavitymd/analysis.py:34-47."""
    re = np.zeros(len(wavevectors))
    im = np.zeros(len(wavevectors))
    for i, k in enumerate(wavevectors):
        kr = np.dot(positions, k)
        re[i] = np.sum(np.cos(kr))
        im[i] = np.sum(np.sin(kr))
    return re + 1j * im


def numpy_field_autocorr(field0, field_t):
    """src/cavitymd/analysis.py:359-364."""
    return np.mean(np.real(field0 * np.conj(field_t)))


def numpy_total_dipole(position, image, charge, box_lengths):
    """compute_total_dipole_moment (reference src/cavitymd/analysis.py:18-31): np.dot(charge, unwrapped)."""
    return np.dot(charge, numpy_unwrap(position, image, box_lengths))


def numpy_cavity_mode(position, image, velocity, mass, typeid, box_lengths, harmonic_energy, kb=3.167e-6):
    """CavityModeTracker.compute_cavity_properties (reference src/cavitymd/analysis.py:1324-1371):
    -> (kinetic, potential, total, temperature) of the first typeid == 2 particle, zeros if none."""
    mask = typeid == 2
    if not np.any(mask):
        return 0.0, 0.0, 0.0, 0.0
    m = mass[mask][0]
    v = velocity[mask][0]
    ke = 0.5 * m * np.sum(v ** 2)
    return ke, harmonic_energy, ke + harmonic_energy, (2.0 / 3.0) * ke / kb
