"""Pick the labelled-array backend: real ``xarray`` when it is installed, else the bundled stand-in."""

try:  # pragma: no cover - xarray is absent from the build image and the GPU box
    import xarray as xr

    HAVE_XARRAY = True
except ModuleNotFoundError:
    from . import xarray_lite as xr

    HAVE_XARRAY = False

__all__ = ["xr", "HAVE_XARRAY"]
