"""ctypes binding of the C ABI in ``include/cse.h``.

``load()`` returns the binding of ``libcse_sm100a.so`` (built in-tree by ``build.py``).  There
is no CPU fallback on this path: if the library is missing or cannot be loaded the import of
any compute entry point raises ``CseLibraryError``.
"""
import ctypes
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))

CSE_OK = 0
CSE_EINVAL, CSE_ECUDA, CSE_EUNSUPPORTED, CSE_EWORKSPACE = -1, -2, -3, -4
ALG_SS, ALG_WIENER, ALG_MMSE, ALG_OMLSA = 0, 1, 2, 3
FLAG_VALID, FLAG_ALIGNED, FLAG_SNR_INF, FLAG_STOI_SHORT = 1, 2, 4, 8

_vp = ctypes.c_void_p
_i = ctypes.c_int
_d = ctypes.c_double
_sz = ctypes.c_size_t

#: every symbol include/cse.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "cse_abi_version": (_i, []),
    "cse_dtype": (_i, []),
    "cse_last_error": (ctypes.c_char_p, []),
    "cse_max_score_length": (_i, [_i]),
    "cse_bins_padded": (_i, [_i]),
    "cse_num_frames": (_i, [_i, _i]),
    "cse_tables_bytes": (_sz, []),
    "cse_tables_init": (_i, [_vp, _vp]),
    "cse_stft_psd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _d, _vp, _vp, _vp]),
    "cse_noise_workspace_bytes": (_sz, [_i, _i, _i]),
    "cse_noise_percentile": (_i, [_vp, _i, _i, _i, _d, _d, _vp, _vp, _sz, _vp]),
    "cse_noise_mintrack": (_i, [_vp, _i, _i, _i, _d, _vp, _vp, _sz, _vp]),
    "cse_gamma": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _d, _d, _vp, _vp]),
    "cse_enhance": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp]),
    "cse_clean_cache_bytes": (_sz, [_i, _i]),
    "cse_clean_workspace_bytes": (_sz, [_i, _i, _i]),
    "cse_prepare_clean": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "cse_score_workspace_bytes": (_sz, [_i, _i, _i]),
    "cse_score": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _sz, _vp]),
    "cse_sweep_workspace_bytes": (_sz, [_i, _i, _i]),
    "cse_enhance_items": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _i, _vp, _i, _i, _i, _vp, _vp]),
    "cse_enhance_list": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _i, _vp, _i, _vp, _i, _vp, _vp]),
    "cse_enhance_groups": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _i, _vp]),
    "cse_gamma_groups": (_i, [_i, _i, _i, _vp, _i, _vp]),
    "cse_score_items": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _sz, _vp]),
    "cse_align_items": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _sz, _vp]),
    "cse_stoi_items": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _sz, _vp]),
    "cse_expand_scores": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp]),
    "cse_select_best": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "cse_debug_special": (_i, [_i, _vp, _vp, _i]),
    "cse_sweep": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp, _i, _i, _vp, _vp, _vp, _i, _vp, _sz, _vp]),
}


#: cse_winner_t (include/cse.h)
WINNER_DTYPE = np.dtype([("index", np.int32), ("lag", np.int32), ("flags", np.int32), ("reserved", np.int32),
                         ("score", np.float64), ("stoi", np.float64), ("pesq", np.float64), ("snr", np.float64)])


class EnhanceGroup(ctypes.Structure):
    """``cse_enhance_group`` (include/cse.h): one noise-PSD group of a grouped enhance launch."""
    _fields_ = [("Y", _vp), ("N", _vp), ("params", _vp), ("out", _vp), ("hop", _i), ("n_params", _i)]


class GammaGroup(ctypes.Structure):
    """``cse_gamma_group`` (include/cse.h)."""
    _fields_ = [("Y", _vp), ("N", _vp), ("G", _vp), ("noise_tv", _i), ("hop", _i), ("noise_mu", _d), ("eps", _d)]


class CseLibraryError(RuntimeError):
    pass


class CseError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"cse error {code}: {message}")
        self.code = code


class CseLibrary:
    """Thin typed wrapper; every compute call raises CseError on a non-zero status."""

    def __init__(self, path):
        if not os.path.exists(path):
            raise CseLibraryError(
                f"{path} not found - build it with `python -m classical_speech_enhancement_b200.build` "
                "(there is no CPU fallback for this path)")
        try:
            self._dll = ctypes.CDLL(path)
        except OSError as e:  # pragma: no cover
            raise CseLibraryError(f"cannot load {path}: {e}") from e
        self.path = path
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(self._dll, name)
            except AttributeError as e:
                raise CseLibraryError(f"{path} does not export {name}") from e
            fn.restype = res
            fn.argtypes = args
        if self._dll.cse_abi_version() != 1:
            raise CseLibraryError("ABI version mismatch")
        self.real_bits = self._dll.cse_dtype()
        self.real = np.float64 if self.real_bits == 64 else np.float32
        self.score_dtype = np.dtype([("stoi", self.real), ("snr", self.real), ("lag", np.int32), ("flags", np.int32)])
        self.winner_dtype = WINNER_DTYPE

    def last_error(self):
        return self._dll.cse_last_error().decode("utf-8", "replace")

    def __getattr__(self, name):
        sym = "cse_" + name
        if sym not in SIGNATURES:
            raise AttributeError(name)
        fn = getattr(self._dll, sym)
        if SIGNATURES[sym][0] is not _i or name in ("abi_version", "dtype", "bins_padded", "num_frames", "max_score_length"):
            self.__dict__[name] = fn                 # next lookup is a plain attribute, not __getattr__
            return fn

        def call(*args):
            rc = fn(*args)
            if rc != CSE_OK:
                raise CseError(rc, self.last_error())
            return rc
        self.__dict__[name] = call
        return call


_cached = {}


def load(fp64=False):
    """The product library (fp32 by default).  Raises CseLibraryError if it is not built."""
    from . import build
    path = build.lib_path(fp64)
    if path not in _cached:
        _cached[path] = CseLibrary(path)
    return _cached[path]


def pack_params(rows):
    """rows: iterable of up to 8 floats -> (n, 8) float64 array laid out as cse_params[]."""
    out = np.zeros((len(rows), 8), dtype=np.float64)
    for i, r in enumerate(rows):
        out[i, :len(r)] = r
    return out
