"""Host-side description of the additional observables g(r) and S(k) (the reference's
``CalculateAdditionalSystemProperties``, BosonsBulk.cpp:474-520, NUBosonsBulkPB.cpp:597-639): plain data for
``tdvmc_observable_desc`` (include/tdvmc_gpu.h).  Nothing here computes an observable."""
from dataclasses import dataclass

import numpy as np


@dataclass
class ObservableSpec:
    gr_count: int            # pairDistribution.grid.count
    gr_spacing: float        # pairDistribution.grid.spacing
    gr_max: float            # pairDistribution.grid.max
    gr_weight: float         # BosonsBulk: DIM / (N - 1) (BosonsBulk.cpp:481); NUBosonsBulkPB: 1 (NUBosonsBulkPB.cpp:611)
    gr_scaling: np.ndarray   # [gr_count] shell volumes (ObservableVsOnGridWithScaling.cpp:19-44)
    shell_ptr: np.ndarray    # [n_shells + 1] int32
    kvec: np.ndarray         # [n_kvec][3], already multiplied by 2 pi / L (BosonsBulk.cpp:127-137)

    @property
    def n_shells(self):
        return len(self.shell_ptr) - 1


def pair_distribution_grid(r_max, n_bins, dim=3):
    """Grid::Init(0, r_max, r_max / n_bins) (Grid.cpp:16-31) and InitScaling (ObservableVsOnGridWithScaling.cpp:19-44).
    ``count`` is the truncated quotient, so it can come out one short of ``n_bins`` (3.5 / (3.5 / 100) -> 99)."""
    spacing = r_max / n_bins
    count = int((r_max - 0.0) / spacing)
    if dim != 3:
        raise ValueError("3-D only")
    vol = np.array([4.0 * np.pi * (spacing * (i + 1)) ** 3.0 / 3.0 for i in range(count)])
    scaling = vol.copy()
    for i in range(count - 1, 0, -1):
        scaling[i] = scaling[i] - scaling[i - 1]
    return count, spacing, scaling


def wave_vectors(shells, lbox, n_shells=50):
    """``shells``: integer wave vectors grouped by norm, the content of kVectors3D.json; the first ``n_shells``
    (numOfkValues = 50, BosonsBulk.cpp:55) are used, each scaled by 2 pi / L (BosonsBulk.cpp:127-137)."""
    ptr = [0]
    rows = []
    for sh in shells[:n_shells]:
        for v in sh:
            rows.append([float(x) * (2 * np.pi / lbox) for x in v])     # kValues[k][kn][a] *= 2 * M_PI / LBOX
        ptr.append(len(rows))
    return np.array(ptr, np.int32), np.array(rows, np.float64).reshape(-1, 3)


def bulk_observables(system_name, n_particles, r_max, gr_bin_count, shells, lbox, n_shells=50):
    count, spacing, scaling = pair_distribution_grid(r_max, gr_bin_count)
    weight = 1.0 / float(n_particles - 1) * 3 if system_name == "BosonsBulk" else 1.0
    ptr, kv = wave_vectors(shells, lbox, n_shells)
    return ObservableSpec(count, spacing, r_max, weight, scaling, ptr, kv)


def from_golden(g):
    """The observable description stored in a tests/golden/*_obs.npz fixture (the reference's own grid and vectors)."""
    sizes = np.asarray(g["k_shell_sizes"]).astype(np.int64)
    ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    return ObservableSpec(int(g["gr_count"]), float(g["gr_spacing"]), float(g["gr_max"]), float(g["gr_weight"]),
                          np.asarray(g["gr_scaling"], np.float64), ptr, np.asarray(g["k_vectors"], np.float64).reshape(-1, 3))
