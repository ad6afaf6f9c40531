"""Vocabulary strings of the hot path, re-declared (not imported) from the reference.

Reference: ``src/xmris/core/config.py`` -- ``XmrisTerm`` ``:9-44``, ``ATTRS`` ``:128-200``, ``DIMS`` ``:229-249``,
``COORDS`` ``:277-283``.  Only the terms the four accessor methods touch are declared.
"""

from __future__ import annotations


class XmrisTerm(str):
    """A ``str`` that also carries ``.unit`` / ``.description`` and a display ``.long_name`` (config.py:9-44)."""

    def __new__(cls, value: str, description: str = "", unit: str = ""):
        obj = str.__new__(cls, value)
        obj.description = description
        obj.unit = unit
        return obj

    @property
    def long_name(self) -> str:
        return self.replace("_", " ").title()


class _Attrs:
    reference_frequency = XmrisTerm("reference_frequency", "Measured Larmor frequency of the target nucleus.", "MHz")
    carrier_ppm = XmrisTerm("carrier_ppm", "Absolute chemical shift at 0 Hz of the baseband signal.", "ppm")
    phase_p0 = XmrisTerm("phase_p0", "Zero-order phase angle applied.", "degrees")
    phase_p1 = XmrisTerm("phase_p1", "First-order phase angle applied.", "degrees")
    phase_pivot = XmrisTerm("phase_pivot", "Coordinate value the first-order phase is anchored at.", "dimension-dependent")
    phase_pivot_coord = XmrisTerm("phase_pivot_coord", "The coordinate dimension in which the phase pivot was defined.")
    apodization_lb = XmrisTerm("apodization_lb", "Line broadening factor applied.", "Hz")
    apodization_gb = XmrisTerm("apodization_gb", "Gaussian broadening factor applied.", "Hz")
    zero_fill_target = XmrisTerm("zero_fill_target", "Total number of points after zero-filling.")
    zero_fill_position = XmrisTerm("zero_fill_position", "Position of padding ('end' or 'symmetric').")
    baseline_method = XmrisTerm("baseline_method", "The algorithm used to estimate and remove the spectral baseline.")
    baseline_lam = XmrisTerm("baseline_lam", "The smoothness penalty (lambda) applied during AsLS baseline correction.")
    baseline_p = XmrisTerm("baseline_p", "The asymmetry parameter applied during AsLS baseline correction.")
    baseline_iter = XmrisTerm("baseline_iter", "The number of sparse solver iterations used to calculate the baseline.")


class _Dims:
    time = XmrisTerm("time", "Time-domain dimension for Free Induction Decay (FID) data.")
    frequency = XmrisTerm("frequency", "Frequency-domain dimension.")
    chemical_shift = XmrisTerm("chemical_shift", "Chemical shift dimension.")
    component = XmrisTerm("component", "Dimension separating real and imaginary parts.")


class _Coords:
    time = XmrisTerm("time", "Time coordinates.", "s")
    frequency = XmrisTerm("frequency", "Frequency coordinates.", "Hz")
    chemical_shift = XmrisTerm("chemical_shift", "Chemical shift coordinates.", "ppm")


ATTRS = _Attrs()
DIMS = _Dims()
COORDS = _Coords()
