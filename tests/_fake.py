"""Test doubles: the analysis classes with the per-frame GPU work replaced by the
CPU oracle, so the host-side loop / sharding / reduction can run without a GPU."""
import numpy as np

from mdhelper_b200.analysis import structure
from oracle import reference_port as rp


class FakeRDF(structure.RadialDistributionFunction):
    def _process(self, frames):
        self._local_counts = np.zeros(self._n_bins, dtype=np.int64)
        for f in frames:
            ts = self._trajectory[int(f)]
            dims = ts.dimensions.copy()
            p1, p2 = self.ag1.positions, self.ag2.positions
            if self._drop_axis is None:
                self._area_or_volume += ts.volume
            else:
                p1[:, self._drop_axis] = p2[:, self._drop_axis] = 0
                dims[self._drop_axis] = dims[:3].max()
                self._area_or_volume += float(
                    np.delete(dims[:3], self._drop_axis).astype(np.float64).prod())
            self._local_counts += rp.radial_histogram(
                p1, p2, self._n_bins, self._range, dims, exclusion=self._exclusion)


class FakeSSF(structure.StructureFactor):
    def _process(self, frames):
        self._local_ssf = np.zeros_like(self.results.ssf)
        offs = np.concatenate(([0], np.cumsum(self._Ns)))
        for f in frames:
            self._trajectory[int(f)]
            pos = np.concatenate([g.positions for g in self._groups]).astype(np.float64)
            for i, (j, k) in enumerate(self.results.pairs):
                if j is None:
                    r = rp.delta_fourier_transform_sum(self._wavevectors, pos)
                    self._local_ssf[i] += (r * r.conj()).real
                    continue
                rj = rp.delta_fourier_transform_sum(self._wavevectors,
                                                    pos[offs[j]:offs[j + 1]])
                if j == k:
                    self._local_ssf[i] += (rj * rj.conj()).real
                else:
                    rk = rp.delta_fourier_transform_sum(self._wavevectors,
                                                        pos[offs[k]:offs[k + 1]])
                    self._local_ssf[i] += 2 * (rj * rk.conj()).real


class FakeISF(structure.IntermediateScatteringFunction):
    """rho(q, t) and the displacement sums from the CPU oracle for this rank's
    wavevector columns; the window logic is the reference's (structure.py:1959-2033)."""

    def _process(self, frames):
        from mdhelper_b200.analysis.base import world
        rank, size = world()
        n_q = len(self._wavenumbers)
        cols = np.array_split(np.arange(n_q), size)[rank]
        self._local_cols = cols
        offs = np.concatenate(([0], np.cumsum(self._Ns)))
        n_rho = 1 if self._mode is None else self._n_groups
        cisf = np.zeros_like(self.results.cisf)
        iisf = np.zeros_like(self.results.iisf) if self._incoherent else None
        self._local = (cisf, iisf)
        if len(cols) == 0:
            return
        wv = self._wavevectors[cols]
        rows = [slice(0, int(self._N))] if self._mode is None else \
            [slice(offs[i], offs[i + 1]) for i in range(self._n_groups)]
        pos, rho = [], []
        for f in frames:
            self._trajectory[int(f)]
            p = np.concatenate([g.positions for g in self._groups]).astype(np.float64)
            pos.append(p)
            rho.append([rp.delta_fourier_transform_sum(wv, p[r]) for r in rows])
        for t in range(len(frames)):
            for lag in range(min(self._n_lags, t + 1)):
                for i, (j, k) in enumerate(self.results.pairs):
                    if j is None:
                        j = k = 0
                    cisf[lag, i, cols] += (rho[t - lag][j] * rho[t][k].conj()).real
                    if j != k:
                        cisf[lag, i, cols] += (rho[t - lag][k] * rho[t][j].conj()).real
                    elif self._incoherent:
                        iisf[lag, j, cols] += rp.delta_fourier_transform_sum(
                            wv, pos[t][rows[j]] - pos[t - lag][rows[j]]).real
