"""xmris_b200 -- B200-native drop-in for xmris's FID->spectrum hot path
(``zero_fill -> apodize_exp -> to_spectrum -> autophase``), see DESIGN.md.

Importing the package registers the ``.xmr`` accessor (on ``xarray`` when installed, and on the bundled
``xarray_lite`` stand-in).  Nothing is computed on the CPU: the data path needs ``libxmris_b200.so`` and a CUDA device.
"""

from .accessor import XmrisB200Accessor, register
from .processing import (apodize_exp, apodize_lg, autophase, baseline_als, fft, fftc, fftshift, ifft, ifftc, ifftshift,  # noqa: F401
                         phase, process_fid, remove_digital_filter, to_complex, to_fid, to_hz, to_ppm, to_real_imag,
                         to_spectrum, zero_fill)
from .vocab import ATTRS, COORDS, DIMS  # noqa: F401
from ._xr import HAVE_XARRAY, xr  # noqa: F401

register()

__version__ = "0.1.0"
__all__ = ["zero_fill", "apodize_exp", "apodize_lg", "to_spectrum", "to_fid", "phase", "autophase", "process_fid", "to_ppm", "to_hz", "fft", "ifft", "fftshift", "ifftshift", "fftc", "ifftc", "remove_digital_filter", "to_real_imag", "to_complex", "baseline_als",
           "ATTRS", "DIMS", "COORDS", "XmrisB200Accessor", "xr"]
