"""The ``.xmr`` accessor for the hot path -- drop-in for the reference's ``XmrisAccessor`` methods
``zero_fill`` / ``apodize_exp`` / ``to_spectrum`` / ``phase`` / ``autophase`` (``src/xmris/core/accessor.py:452-683``),
same names, argument order and defaults (checked by the reference's ``tests/test_core.py:509-552``), plus the "next"
rows ``to_fid`` / ``apodize_lg`` / ``to_ppm`` / ``to_hz`` / ``remove_digital_filter`` and the fused entry point
``process_fid``.

Registered on the bundled ``xarray_lite.DataArray`` always, and on real ``xarray.DataArray`` when xarray imports
(registering under ``"xmr"`` overrides the reference's accessor if both packages are imported).
"""

from __future__ import annotations

from . import processing as P
from ._xr import HAVE_XARRAY, xr
from .processing import _check_dims  # noqa: F401  (the reference re-exports it here; tests/test_core.py:49)
from .vocab import DIMS


class XmrisB200Accessor:
    def __init__(self, xarray_obj):
        self._obj = xarray_obj

    # --- coordinate systems (accessor.py:329-366; host metadata only) ---
    def to_ppm(self, dim: str = DIMS.frequency):
        return P.to_ppm(self._obj, dim=dim)

    def to_hz(self, dim: str = DIMS.chemical_shift):
        return P.to_hz(self._obj, dim=dim)

    # --- Fourier mixin (accessor.py:369-447) ---
    def fft(self, dim=DIMS.time, out_dim=None):
        return P.fft(self._obj, dim=dim, out_dim=out_dim)

    def ifft(self, dim=DIMS.frequency, out_dim=None):
        return P.ifft(self._obj, dim=dim, out_dim=out_dim)

    def fftshift(self, dim):
        return P.fftshift(self._obj, dim=dim)

    def ifftshift(self, dim):
        return P.ifftshift(self._obj, dim=dim)

    def fftc(self, dim=DIMS.time, out_dim=None):
        return P.fftc(self._obj, dim=dim, out_dim=out_dim)

    def ifftc(self, dim=DIMS.frequency, out_dim=None):
        return P.ifftc(self._obj, dim=dim, out_dim=out_dim)

    # --- processing (accessor.py:452-550) ---
    def apodize_exp(self, dim: str = DIMS.time, lb: float = 1.0):
        return P.apodize_exp(self._obj, dim=dim, lb=lb)

    def apodize_lg(self, dim: str = DIMS.time, lb: float = 1.0, gb: float = 1.0):
        return P.apodize_lg(self._obj, dim=dim, lb=lb, gb=gb)

    def to_spectrum(self, dim: str = DIMS.time, out_dim: str = DIMS.frequency):
        return P.to_spectrum(self._obj, dim=dim, out_dim=out_dim)

    def to_fid(self, dim: str = DIMS.frequency, out_dim: str = DIMS.time):
        return P.to_fid(self._obj, dim=dim, out_dim=out_dim)

    def zero_fill(self, dim: str = DIMS.time, target_points: int = 1024, position: str = "end"):
        return P.zero_fill(self._obj, dim=dim, target_points=target_points, position=position)

    # --- baseline (accessor.py:552-597) ---
    def baseline_als(self, dim: str = DIMS.frequency, lam: float = 1e5, p: float = 0.001, n_iter: int = 10):
        return P.baseline_als(self._obj, dim=dim, lam=lam, p=p, n_iter=n_iter)

    # --- phasing (accessor.py:599-683) ---
    def phase(self, dim: str = DIMS.frequency, p0: float = 0.0, p1: float = 0.0, pivot: float = None):
        return P.phase(self._obj, dim=dim, p0=p0, p1=p1, pivot=pivot)

    def autophase(self, dim: str = DIMS.frequency, method: str = "acme", peak_width: int = 100, lb: float = 0.0,
                  temp_time_dim: str = DIMS.time, **kwargs):
        # NB the accessor default peak_width=100 differs from the function default 0.5 (accessor.py:634 vs
        # phasing.py:166); mode / target_coord / p0_only travel in **kwargs (accessor.py:637, 682).
        return P.autophase(self._obj, dim=dim, method=method, peak_width=peak_width, lb=lb,
                           temp_time_dim=temp_time_dim, **kwargs)

    # --- vendor specific (accessor.py:829-859) ---
    def remove_digital_filter(self, group_delay: float, dim: str = "time", keep_length: bool = True):
        return P.remove_digital_filter(self._obj, group_delay=group_delay, dim=dim, keep_length=keep_length)

    # --- utility / formatting (accessor.py:863-878; host data re-labelling only) ---
    def to_real_imag(self, dim: str = DIMS.component, coords: tuple = ("real", "imag")):
        return P.to_real_imag(self._obj, dim=dim, coords=coords)

    def to_complex(self, dim: str = DIMS.component, coords: tuple = ("real", "imag")):
        return P.to_complex(self._obj, dim=dim, coords=coords)

    # --- fused chain (B200 extension) ---
    def process_fid(self, dim: str = DIMS.time, out_dim: str = DIMS.frequency, target_points: int | None = None,
                    position: str = "end", lb: float | None = None, autophase_kwargs: dict | None = None,
                    baseline_kwargs: dict | None = None, gb: float | None = None):
        return P.process_fid(self._obj, dim=dim, out_dim=out_dim, target_points=target_points, position=position,
                             lb=lb, autophase_kwargs=autophase_kwargs, baseline_kwargs=baseline_kwargs, gb=gb)

    # --- everything this package does not rebuild (plot, widget, fit_amares, ...) -------------------------------------
    _fallback_cls = None   # the accessor class that owned ".xmr" before register() (the reference's, when it is installed)

    def __getattr__(self, name):
        # only reached for attributes this class does not define: hand them to the accessor we displaced, so that importing
        # xmris_b200 next to the reference does not make .xmr.plot / .xmr.widget / .xmr.fit_amares disappear
        fb = type(self)._fallback_cls
        if fb is not None and not name.startswith("__"):
            return getattr(fb(self._obj), name)
        raise AttributeError(f"{type(self).__name__!s} has no attribute {name!r} (not on the FID->spectrum hot path; "
                             "install the reference package for plotting / fitting)")


def register():
    """(Re-)register the accessor under ``.xmr`` on every available DataArray type."""
    import warnings

    from . import xarray_lite

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        xarray_lite.register_dataarray_accessor("xmr")(XmrisB200Accessor)
        if HAVE_XARRAY:  # pragma: no cover
            prev = xr.DataArray.__dict__.get("xmr")
            prev_cls = getattr(prev, "_accessor", None)
            if prev_cls is not None and prev_cls is not XmrisB200Accessor:
                XmrisB200Accessor._fallback_cls = prev_cls      # keep the displaced accessor reachable (see __getattr__)
            xr.register_dataarray_accessor("xmr")(XmrisB200Accessor)
    return XmrisB200Accessor
