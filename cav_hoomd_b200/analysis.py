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
