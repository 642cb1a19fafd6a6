"""Deterministic synthetic inputs in HOOMD's array layouts.

The benchmark configs of BASELINE.json name "synthetic 1M-particle charged LJ box"; the real
input of config 1 (examples/init-0.gsd) is not in the reference mount.  This module builds the
stand-in: an O/N charged box with one photon particle of type 'L' appended last, exactly the
way the reference's driver script appends it (examples/05_advanced_run.py:462-505).

Layouts (what HOOMD's ParticleData hands a ForceCompute on the device):
    pos    float64[N,4]  x, y, z, w  -- w carries the type id in its LOW 32 BITS
    vel    float64[N,4]  vx, vy, vz, mass
    charge float64[N]
    image  int32[N,3]
    force  float64[N,4]  fx, fy, fz, potential energy   (output)
"""
from __future__ import annotations

import dataclasses

import numpy as np

# a.u. constants the reference uses (src/cavitymd/utils.py:12-18)
HARTREE_TO_CM_MINUS1 = 219474.63
KB_HARTREE_PER_K = 3.167e-6

TYPES = ("O", "N", "L")
L_TYPEID = 2
MASS_O = 29166.0
MASS_N = 25520.0
NUMBER_DENSITY = 5.5e-3  # particles / Bohr^3
KT_100K = 100.0 * KB_HARTREE_PER_K  # 3.167e-4 Ha
DT_1FS = 41.3414  # a.u.
TAU_5PS = 206706.9  # a.u.


def typeid_to_w(typeid: np.ndarray) -> np.ndarray:
    """Pack int32 type ids into the low 32 bits of a float64 (HOOMD's __int_as_scalar)."""
    w = np.zeros(len(typeid), dtype=np.uint64)
    w |= np.asarray(typeid, dtype=np.int32).view(np.uint32).astype(np.uint64)
    return w.view(np.float64)


def w_to_typeid(w: np.ndarray) -> np.ndarray:
    """Low 32 bits of pos.w as int32 (HOOMD's __scalar_as_int)."""
    return (np.ascontiguousarray(w).view(np.uint64) & np.uint64(0xFFFFFFFF)).astype(np.uint32).view(np.int32)


@dataclasses.dataclass
class System:
    """One particle system in HOOMD's layouts (host copies)."""

    pos: np.ndarray
    vel: np.ndarray
    charge: np.ndarray
    image: np.ndarray
    box: tuple
    L_typeid: int = L_TYPEID
    types: tuple = TYPES

    @property
    def N(self) -> int:
        return int(self.pos.shape[0])

    @property
    def typeid(self) -> np.ndarray:
        return w_to_typeid(self.pos[:, 3])

    def copy(self) -> "System":
        return System(self.pos.copy(), self.vel.copy(), self.charge.copy(), self.image.copy(),
                      self.box, self.L_typeid, self.types)


def make_system(n_mol: int, replica: int = 0, *, photon: str = "last", images: bool = True,
                charges: str = "neutral", omegac: float = 0.01, kT: float = KT_100K,
                phmass: float = 1.0) -> System:
    """Appendix D recipe.  `photon` in {"last","first","middle","absent","duplicated"};
    `charges` in {"neutral","zero","nonneutral"}."""
    rng = np.random.Generator(np.random.PCG64(20261018 + replica))
    L = (max(n_mol, 1) / NUMBER_DENSITY) ** (1.0 / 3.0)
    # GSD stores float32 positions: round, then widen
    xyz = rng.uniform(-L / 2, L / 2, size=(n_mol, 3)).astype(np.float32).astype(np.float64)
    img = rng.integers(-1, 2, size=(n_mol, 3), dtype=np.int32) if images else np.zeros((n_mol, 3), np.int32)
    idx = np.arange(n_mol)
    typeid = (idx & 1).astype(np.int32)  # O even, N odd
    if charges == "neutral":
        q = np.where(idx & 1, -0.5, 0.5)
    elif charges == "zero":
        q = np.zeros(n_mol)
    elif charges == "nonneutral":
        q = rng.uniform(-1, 1, size=n_mol).astype(np.float32).astype(np.float64) + 0.25
    else:
        raise ValueError(charges)
    mass = np.where(idx & 1, MASS_N, MASS_O)
    v = rng.standard_normal((n_mol, 3)) * np.sqrt(kT / mass)[:, None]

    K = phmass * omegac * omegac
    qph = rng.standard_normal(3) * np.sqrt(kT / K)
    vph = rng.standard_normal(3) * np.sqrt(kT / phmass)

    def rows(xyz_, tid, q_, img_, v_, m_):
        pos_ = np.empty((len(tid), 4))
        pos_[:, :3] = xyz_
        pos_[:, 3] = typeid_to_w(tid)
        vel_ = np.empty((len(tid), 4))
        vel_[:, :3] = v_
        vel_[:, 3] = m_
        return pos_, vel_, np.asarray(q_, dtype=np.float64), np.asarray(img_, dtype=np.int32)

    mol = rows(xyz, typeid, q, img, v, mass)
    ph = rows(qph[None, :], np.array([L_TYPEID], np.int32), [0.0], np.zeros((1, 3), np.int32),
              vph[None, :], [phmass])

    def cat(parts):
        return tuple(np.ascontiguousarray(np.concatenate([p[i] for p in parts])) for i in range(4))

    if photon == "last":
        pos, vel, charge, image = cat([mol, ph])
    elif photon == "first":
        pos, vel, charge, image = cat([ph, mol])
    elif photon == "middle":
        h = n_mol // 2
        a = tuple(x[:h] for x in mol)
        b = tuple(x[h:] for x in mol)
        pos, vel, charge, image = cat([a, ph, b])
    elif photon == "absent":
        pos, vel, charge, image = cat([mol])
    elif photon == "duplicated":
        # a second 'L' particle WITH a charge: the reference skips only the first one in the
        # dipole sum (src/CavityForceCompute.cc:122) and gives the second zero force (:190)
        ph2 = rows(qph[None, :] * 0.5, np.array([L_TYPEID], np.int32), [0.75],
                   np.ones((1, 3), np.int32), vph[None, :], [phmass])
        h = n_mol // 3
        a = tuple(x[:h] for x in mol)
        b = tuple(x[h:] for x in mol)
        pos, vel, charge, image = cat([a, ph, b, ph2])
    else:
        raise ValueError(photon)
    return System(pos, vel, charge, image, (L, L, L))


def molecular_group(system: System) -> np.ndarray:
    """Indices of the thermostatted group filter.Type(['O','N']) (everything but 'L')."""
    return np.nonzero(system.typeid != system.L_typeid)[0].astype(np.uint32)


def fibonacci_sphere(samples: int) -> np.ndarray:
    """K unit vectors, the construction of src/cavitymd/analysis.py:50-66 (vectorised)."""
    i = np.arange(samples, dtype=np.float64)
    phi = np.pi * (3.0 - np.sqrt(5.0))
    y = 1 - (i / float(samples - 1)) * 2
    radius = np.sqrt(1 - y * y)
    theta = phi * i
    return np.stack([np.cos(theta) * radius, y, np.sin(theta) * radius], axis=1)


def random_walk_frames(system: System, T: int, sigma: float = 0.05, seed: int = 7) -> np.ndarray:
    """T wrapped snapshots float64[T,N,3] from a random walk of the box (Appendix D)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    L = np.asarray(system.box)
    x = system.pos[:, :3].copy()
    out = np.empty((T,) + x.shape)
    for t in range(T):
        out[t] = x
        x = x + rng.standard_normal(x.shape) * sigma
        x = (x + L / 2) % L - L / 2
    return out
