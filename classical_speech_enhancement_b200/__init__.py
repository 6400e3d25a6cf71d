"""B200-native enhancement-and-scoring sweep, drop-in for the hot path of
Katja39/Classical_Speech_Enhancement (see DESIGN.md / INTEGRATION.md).

Importing the package is cheap; the native library ``libcse_sm100a.so`` is loaded on first use
and its absence is an error (no CPU fallback).
"""
from .parameter_ranges import (param_ranges_mmse, param_ranges_omlsa, param_ranges_ss,  # noqa: F401
                               param_ranges_wiener)

__all__ = ["param_ranges_ss", "param_ranges_mmse", "param_ranges_wiener", "param_ranges_omlsa"]
