"""algebraist_b200: B200-native S_n fast Fourier transform.

Drop-in for the hot path of dashstander/algebraist (`sn_fft`, `sn_ifft`, plus
`sn_fourier_decomposition`, `calc_power` and the dense cross-checks `slow_sn_ft`, `slow_sn_ift`); mirrors
`/root/reference/src/algebraist/__init__.py:16-19`."""
from . import _lib
from .fourier import (bandlimited_partitions, calc_power, sn_convolve, sn_fft, sn_fft_bandlimited, sn_fourier_decomposition,
                      sn_ifft, sn_ifft_bandlimited, slow_sn_ft, slow_sn_ift)
from .group import all_permutations, permutation_index, rho, sn_translate
from .host_pipeline import stream_batches

__all__ = ["sn_fft", "sn_ifft", "sn_fourier_decomposition", "calc_power", "slow_sn_ft", "slow_sn_ift",
           "sn_convolve", "stream_batches", "sn_fft_bandlimited", "sn_ifft_bandlimited", "bandlimited_partitions", "sn_translate", "rho", "permutation_index", "all_permutations"]
