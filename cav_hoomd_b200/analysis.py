"""F(k,t) analysis, B200 build -- host-side mirror of the density-field part of the reference
src/cavitymd/analysis.py (compute_density_field :34-47, generate_fibonacci_sphere :50-66,
FieldAutocorrelationTracker :260-418).  The K x N phase sum runs in one batched kernel
(cavb200_rhok) over T frames instead of K NumPy passes per step."""
from __future__ import annotations

import numpy as np

from . import capi
from .synth import fibonacci_sphere


def generate_fibonacci_sphere(samples=100):
    """Unit vectors on a sphere (reference analysis.py:50-66)."""
    return fibonacci_sphere(samples)


class DensityField:
    """rho[t][k] = sum_j exp(i k . r_j(t)) on the device."""

    def __init__(self, wavevectors, handle: capi.Handle | None = None, device: int = 0):
        self.wavevectors = np.ascontiguousarray(wavevectors, dtype=np.float64)
        self.K = self.wavevectors.shape[0]
        self._handle = handle or capi.Handle(device)
        self._d_k = capi.DeviceArray.from_numpy(self.wavevectors)

    def compute_frames(self, d_pos, stride: int, frame_stride: int, N: int, T: int, stream=None):
        """Device positions (T frames) -> device rho (T, K, 2)."""
        d_rho = capi.DeviceArray((T, self.K, 2), np.float64)
        self._handle.rhok(d_pos, stride, frame_stride, N, T, self._d_k, self.K, d_rho, stream)
        return d_rho

    def compute_density_field(self, positions: np.ndarray) -> np.ndarray:
        """compute_density_field(snapshot, wavevectors) for host positions [N,3] or [T,N,3]
        (reference analysis.py:34-47) -> complex rho[K] or [T,K]."""
        if np.asarray(positions).dtype == np.float32 and np.asarray(positions).shape[-1] == 3:
            # GSD trajectories: float32 xyz goes to the device as it is (12 B/particle/frame) and is widened there
            p = np.ascontiguousarray(positions, dtype=np.float32)
            single = p.ndim == 2
            if single:
                p = p[None]
            T, N, _ = p.shape
            d_pos = capi.DeviceArray.from_numpy(p)
            d_rho = capi.DeviceArray((T, self.K, 2), np.float64)
            self._handle.rhok_f32(d_pos, 3 * N, N, T, self._d_k, self.K, d_rho)
            r = d_rho.numpy()
            rho = r[..., 0] + 1j * r[..., 1]
            return rho[0] if single else rho
        p = np.ascontiguousarray(positions, dtype=np.float64)
        single = p.ndim == 2
        if single:
            p = p[None]
        T, N, stride = p.shape
        d_pos = capi.DeviceArray.from_numpy(p)
        r = self.compute_frames(d_pos, stride, N * stride, N, T).numpy()
        rho = r[..., 0] + 1j * r[..., 1]
        return rho[0] if single else rho

    def autocorrelation(self, d_rho, T: int, n_origins: int, n_lags: int, stream=None) -> np.ndarray:
        """F[o][l] = mean_k Re(rho[o] conj(rho[o+l])) (compute_field_autocorr, analysis.py:359-364)."""
        d_out = capi.DeviceArray((n_origins, n_lags), np.float64)
        self._handle.fkt(d_rho, T, self.K, n_origins, n_lags, d_out, stream)
        return d_out.numpy()


class FieldAutocorrelationTracker:
    """density_correlation tracker: references every `reference_interval_steps` up to
    `max_references`, autocorrelation of each reference with the current field
    (reference analysis.py:260-418; text output and hoomd.custom.Action plumbing left to the caller)."""

    def __init__(self, state, kmag=1.0, num_wavevectors=50, reference_interval_steps=10000, max_references=10,
                 handle: capi.Handle | None = None):
        self.state = state
        self.kmag = kmag
        self.num_wavevectors = num_wavevectors
        self.wavevectors = generate_fibonacci_sphere(num_wavevectors) * kmag  # analysis.py:308-310
        self.field = DensityField(self.wavevectors, handle, state.device)
        self.reference_interval_steps = reference_interval_steps
        self.max_references = max_references
        self.references = []
        self.last_reference_step = 0
        self.current_autocorr_value = 0.0
        self._add_reference(0)

    def _current_field(self):
        s = self.state
        r = self.field.compute_frames(s.pos, 4, 4 * s.N, s.N, 1).numpy()[0]
        return r[:, 0] + 1j * r[:, 1]

    def _add_reference(self, timestep):
        self.references.append({"number": len(self.references), "timestep": timestep, "field": self._current_field()})

    @staticmethod
    def compute_field_autocorr(field0, field_t):
        return float(np.mean(np.real(field0 * np.conj(field_t))))  # analysis.py:359-364

    def act(self, timestep):
        """-> list of (reference number, lag in steps, autocorrelation) (reference act, :380-414)."""
        if timestep == 0:
            return []
        cur = self._current_field()
        out = []
        for ref in self.references:
            v = self.compute_field_autocorr(ref["field"], cur)
            if ref["number"] == 0:
                self.current_autocorr_value = v
            out.append((ref["number"], timestep - ref["timestep"], v))
        if len(self.references) < self.max_references and timestep - self.last_reference_step >= self.reference_interval_steps:
            self._add_reference(timestep)
            self.last_reference_step = timestep
        return out


# ---------------------------------------------------------------------------------------------
# device-side scalar trackers
# ---------------------------------------------------------------------------------------------
KB_HARTREE_PER_K = 3.167e-6  # reference src/cavitymd/utils.py:13


class StepRecorder:
    """One 128-byte record per step, appended on the device (cavb200_track_record) from what the force and
    thermostat kernels left there.  The reference's trackers take sim.state.cpu_local_snapshot every step
    (analysis.py:188,234,1327) for the same scalars; here nothing leaves the device until `read`."""

    def __init__(self, state, handle: capi.Handle, capacity: int = 16384):
        self.state, self.handle, self.capacity = state, handle, capacity
        handle.track_open(capacity)
        self._seen = 0

    def set_reference(self, stream=None):
        self.handle.track_set_reference(stream)

    def act(self, timestep: int, stream=None):
        self.handle.track_record(timestep, self.state.vel, self.state.N, stream)

    def read(self, stream=None):
        """Records appended since the previous read, oldest first (float64[n, 16]; capi.Handle.TRACK_FIELDS)."""
        _, total = self.handle.track_read(0, stream)
        new = min(total - self._seen, self.capacity)
        rec, total = self.handle.track_read(new, stream)
        self._seen = total
        return rec


class DipoleAutocorrelation:
    """C(t) = d(ref) . d(t) (AutocorrelationTracker('dipole'), reference analysis.py:152-258,1424-1446): a new
    reference when an output lands on a multiple of 10000 steps (:247-249); same text format (:192-197,240-242)."""

    def __init__(self, recorder: StepRecorder, dt: float, output_prefix="dipole_autocorr", output_period_steps=1000,
                 write_files=True):
        self.rec, self.dt = recorder, dt
        self.output_prefix, self.output_period_steps, self.write_files = output_prefix, output_period_steps, write_files
        self.output_file_number = 0
        self.last_output_step = 0
        self.current_autocorr_value = None
        self.series = []  # (reference number, timestep, C)
        recorder.set_reference()
        recorder.act(0)
        r = recorder.read()[-1]
        self.current_autocorr_value = float(r[10])
        self._header(0, r[10])

    def _path(self):
        return f"{self.output_prefix}_{self.output_file_number}.txt"

    def _time_ps(self, timestep):
        return self.dt * timestep * 2.418884e-5  # PhysicalConstants.TIME_PS_CONVERSION (reference utils.py:18)

    def _header(self, timestep, c0):
        self.series.append((self.output_file_number, int(timestep), float(c0)))
        if not self.write_files:
            return
        with open(self._path(), "w") as f:
            f.write("# Dipole autocorrelation data\n")
            f.write(f"# Reference number: {self.output_file_number}\n")
            f.write(f"# Output period: {self.output_period_steps} steps\n")
            f.write("# timestep t(ps) C(t)\n")
            f.write(f"{int(timestep)} {self._time_ps(timestep):.6f} {c0:.6f}\n")

    def act(self, timestep: int, stream=None):
        """Per step: one asynchronous device record.  Only an output step reads anything back."""
        if timestep == 0:
            return
        self.rec.act(timestep, stream)
        if timestep - self.last_output_step >= self.output_period_steps:
            r = self.rec.read(stream)[-1]
            self.current_autocorr_value = float(r[10])
            self.series.append((self.output_file_number, timestep, float(r[10])))
            if self.write_files:
                with open(self._path(), "a") as f:
                    f.write(f"{timestep} {self._time_ps(timestep):.6f} {r[10]:.6f}\n")
            self.last_output_step = timestep
            if timestep % 10000 == 0:
                self.output_file_number += 1
                self.rec.set_reference(stream)
                self._header(timestep, float(np.dot(r[1:4], r[1:4])))

    @property
    def current_autocorr(self):
        return self.current_autocorr_value if self.current_autocorr_value is not None else 0.0


class CavityModeTracker:
    """Cavity-mode kinetic / harmonic / total energy and temperature (reference analysis.py:1285-1417) from
    the newest device record instead of a snapshot."""

    def __init__(self, recorder: StepRecorder):
        self.rec = recorder
        self._last = np.zeros(16)

    def refresh(self, stream=None):
        r, _ = self.rec.handle.track_read(1, stream)
        if len(r):
            self._last = r[-1]
        return self

    @property
    def cavity_kinetic_energy(self):
        return float(self._last[11])

    @property
    def cavity_potential_energy_harmonic(self):
        return float(self._last[7])

    @property
    def cavity_total_energy(self):
        return float(self._last[11] + self._last[7])

    @property
    def cavity_temperature(self):
        return (2.0 / 3.0) * float(self._last[11]) / KB_HARTREE_PER_K  # :1367
