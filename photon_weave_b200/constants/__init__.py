"""Physical constants and temporal-profile functions (reference: photon_weave/constants)."""

import numpy as np

C0 = 299792458  # m/s


def gaussian(t: float, t_a: float, omega: float, mu: float = 0, sigma: float = 1) -> float:
    norm = 1 / (np.sqrt(sigma * np.pi))
    return norm * np.exp(-((t - t_a - mu) ** 2) / (2 * sigma**2))
