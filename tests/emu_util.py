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

    def noise_percentile(self, P, n_fft, percentile, eps):
        U, nf, nbp = P.shape
        N = np.zeros((U, nbp), dtype=self.real)
        ws = np.zeros(self.lib.noise_workspace_bytes(U, nf, n_fft), dtype=np.uint8)
        self.lib.noise_percentile(ptr(P), U, nf, n_fft, percentile, eps, ptr(N), ptr(ws), ws.nbytes, None)
        return N

    def noise_mintrack(self, P, n_fft, eps):
        U, nf, nbp = P.shape
        N = np.zeros((U, nf, nbp), dtype=self.real)
        self.lib.noise_mintrack(ptr(P), U, nf, n_fft, eps, ptr(N), None, 0, None)
        return N

    def prepare_clean(self, clean):
        clean = np.ascontiguousarray(np.atleast_2d(clean), dtype=self.real)
        U, L = clean.shape
        rec = self.lib.clean_cache_bytes(L, 16000)
        cache = np.zeros(U * rec, dtype=np.uint8)
        ws = np.zeros(self.lib.clean_workspace_bytes(U, L, 16000), dtype=np.uint8)
        self.lib.prepare_clean(ptr(self.tables), ptr(clean), U, L, 16000, ptr(cache), ptr(ws), ws.nbytes, None)
        return clean, cache

    def score(self, wav, clean, cache, finalize=True):
        """wav [U][C][L] -> structured scores [U][C]."""
        wav = np.ascontiguousarray(wav, dtype=self.real)
        U, C, L = wav.shape
        scores = np.zeros(U * C, dtype=self.lib.score_dtype)
        ws = np.zeros(self.lib.score_workspace_bytes(U * C, L, 16000), dtype=np.uint8)
        self.lib.score(ptr(self.tables), ptr(wav), U, C, L, 16000, ptr(clean), ptr(cache), int(finalize),
                       ptr(scores), ptr(ws), ws.nbytes, None)
        return scores.reshape(U, C)

    def sweep(self, alg, Y, N, L, n_fft, hop, rows, clean, cache, chunk=3):
        U = Y.shape[0]
        params = pack_params(rows)
        N = np.ascontiguousarray(N, dtype=self.real)
        scores = np.zeros(U * len(rows), dtype=self.lib.score_dtype)
        ws = np.zeros(self.lib.sweep_workspace_bytes(chunk, L, 16000), dtype=np.uint8)
        self.lib.sweep(ptr(self.tables), alg, ptr(Y), ptr(N), int(N.ndim == 3), U, L, n_fft, hop, ptr(params),
                       len(rows), 16000, ptr(clean), ptr(cache), ptr(scores), chunk, ptr(ws), ws.nbytes, None)
        return scores.reshape(U, len(rows))


class NumpyBackend:
    """Engine backend over host memory for the thread-emulated library (tests only)."""

    def empty(self, shape, dtype):
        return np.empty(shape, dtype=dtype)

    def zeros(self, shape, dtype):
        return np.zeros(shape, dtype=dtype)

    def from_host(self, arr, pinned=False):
        return np.array(arr, copy=True, order="C")

    def to_host(self, buf):
        return np.array(buf, copy=True)

    def ptr(self, buf):
        return ptr(buf)

    def ptr_at(self, buf, byte_offset):
        return ctypes.c_void_p(buf.ctypes.data + int(byte_offset))

    def stream(self):
        return None

    def synchronize(self):
        pass

    def view_bytes_as(self, buf, dtype, tag=None):
        return buf.view(dtype)


def use_emulated_runtime(fp64=False):
    """Point SweepEngine's default runtime at the emulation build (CPU tests of host logic)."""
    from classical_speech_enhancement_b200 import engine
    engine.configure_runtime(lib=emu_lib(fp64), backend_factory=NumpyBackend)


def use_product_runtime():
    from classical_speech_enhancement_b200 import engine
    engine.configure_runtime(None, None)
