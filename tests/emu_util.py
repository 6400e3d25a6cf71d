"""numpy-side driver of the C ABI for the CPU thread-emulation build (tests only).

Same calls the product makes through ``classical_speech_enhancement_b200._lib.CseLibrary``, with
numpy arrays standing in for device tensors (under emulation "device" memory is host memory).
"""
import ctypes

import numpy as np

from classical_speech_enhancement_b200 import build
from classical_speech_enhancement_b200._lib import CseLibrary, pack_params  # noqa: F401

_libs = {}


def emu_lib(fp64=False):
    if fp64 not in _libs:
        _libs[fp64] = CseLibrary(build.build_emu(fp64=fp64))
    return _libs[fp64]


def ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class Emu:
    def __init__(self, fp64=False):
        self.lib = emu_lib(fp64)
        self.real = self.lib.real
        self.cplx = np.complex128 if fp64 else np.complex64
        self.tables = np.zeros(self.lib.tables_bytes(), dtype=np.uint8)
        self.lib.tables_init(ptr(self.tables), None)

    def stft_psd(self, wav, n_fft, hop, minus=None, floor=0.0):
        wav = np.ascontiguousarray(np.atleast_2d(wav), dtype=self.real)
        if minus is not None:
            minus = np.ascontiguousarray(np.atleast_2d(minus), dtype=self.real)
        U, L = wav.shape
        nf, nbp = self.lib.num_frames(L, hop), self.lib.bins_padded(n_fft)
        Y = np.zeros((U, nf, nbp), dtype=self.cplx)
        P = np.zeros((U, nf, nbp), dtype=self.real)
        self.lib.stft_psd(ptr(self.tables), ptr(wav), ptr(minus), U, L, n_fft, hop, floor, ptr(Y), ptr(P), None)
        return Y, P

    def enhance(self, alg, Y, N, L, n_fft, hop, rows):
        U = Y.shape[0]
        params = pack_params(rows)
        N = np.ascontiguousarray(N, dtype=self.real)
        out = np.zeros((U * len(rows), L), dtype=self.real)
        self.lib.enhance(ptr(self.tables), alg, ptr(Y), ptr(N), int(N.ndim == 3), U, L, n_fft, hop,
                         ptr(params), len(rows), ptr(out), None)
        return out.reshape(U, len(rows), L)

    def layout_noise(self, N_oracle, n_fft):
        """oracle (bins, frames) or (bins, 1) -> library layout [1][frames][nbp] / [1][nbp]."""
        nbp = self.lib.bins_padded(n_fft)
        nb = n_fft // 2 + 1
        if N_oracle.shape[1] == 1:
            out = np.zeros((1, nbp), dtype=self.real)
            out[0, :nb] = N_oracle[:, 0]
        else:
            out = np.zeros((1, N_oracle.shape[1], nbp), dtype=self.real)
            out[0, :, :nb] = N_oracle.T
        return out
