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
