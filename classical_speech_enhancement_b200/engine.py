"""Batched sweep engine: the host side of the (utterance x grid-point) batch.

Owns the device-resident state of one batch of equal-length utterance pairs - waveforms,
clean-side scoring caches, one STFT per (n_fft, hop), one noise PSD per
(n_fft, hop, method, percentile, eps) - and drives the C ABI (``include/cse.h``).  The
reference recomputes all of that for every candidate (``Code/wiener_filter.py:35-46`` +
``Code/noise_estimation.py:184-188``); here it is computed once and shared by the hundreds of
candidates that differ only in gain parameters.

Array storage goes through a small backend so that the same engine logic runs on CUDA tensors
(product: :class:`TorchCudaBackend`, PyTorch is only the buffer carrier) and, in the CPU test
suite, on numpy arrays against the thread-emulated kernels.  There is no CPU compute path in
the product: constructing :class:`SweepEngine` without a CUDA device raises.
"""
import ctypes
import os
from collections import OrderedDict

import numpy as np

from . import _lib
from .grid import ALGORITHM_IDS, alg_eps, grid_points, plan

SR = 16000
# Candidates per launch: a common multiple of the resident CTAs of every hot kernel on 148 SMs (2, 3, 4 per SM ->
# 296, 444, 592), so no launch ends in a partial wave; the candidate-waveform workspace is capped at 8 GiB.
DEFAULT_CHUNK_ITEMS = 14208
MAX_WAV_WORKSPACE_BYTES = 8 << 30

#: runtime used when an engine is constructed without explicit lib/backend.  The product never
#: changes it (-> libcse_sm100a.so + CUDA tensors); the CPU test-suite points it at the
#: thread-emulated build to exercise the host logic without a GPU.
_runtime = {"lib": None, "backend_factory": None, "reuse_result_buffers": False,
            # share the candidate-invariant front of the Wiener / MMSE / Log-MMSE gain rules (cse_gamma) across a group
            "gamma": os.environ.get("CSE_GAMMA", "1") != "0",
            # a single pair: all groups of a grid scored by one align + one STOI launch (sweep_device)
            # and the groups of one kernel instantiation enhanced by one launch (cse_enhance_groups)
            "fuse_single": os.environ.get("CSE_FUSE_SINGLE", "1") != "0"}
_PLAN_CACHE = OrderedDict()      # (algorithm, grid digest, frame-count signature) -> host-side launch plan, LRU
_PLAN_CACHE_MAX = 32
_PINNED_POOL = {}
_TABLES_CACHE = {}               # (library, device) -> constant tables buffer (twiddles, windows, resampler taps)


class GridPoints(list):
    """A grid's point list that remembers its content digest (see :func:`points_digest`); what
    ``sweep.cached_points`` hands out so that repeated sweeps of the same grid skip re-hashing it."""
    __slots__ = ("_cse_digest", "_cse_shapes")


def points_shapes(points):
    """The distinct (n_fft, hop_length, noise_method) triples of a grid (what validation and the frame-count
    signature of a plan depend on) - remembered on a :class:`GridPoints`."""
    sh = getattr(points, "_cse_shapes", None)
    if sh is None:
        sh = sorted({(int(p["n_fft"]), int(p["hop_length"]), p.get("noise_method")) for p in points}, key=str)
        if isinstance(points, GridPoints):
            points._cse_shapes = sh
    return sh


def points_digest(points):
    """Content hash of a grid (order-sensitive: the selection scan depends on the order)."""
    d = getattr(points, "_cse_digest", None)
    if d is None:
        d = hash(tuple(tuple(p.items()) for p in points))
        if isinstance(points, GridPoints):
            points._cse_digest = d
    return d


def configure_runtime(lib=None, backend_factory=None, reuse_result_buffers=None):
    _runtime["lib"] = lib
    _runtime["backend_factory"] = backend_factory
    if reuse_result_buffers is not None:
        _runtime["reuse_result_buffers"] = bool(reuse_result_buffers)


def reuse_result_buffers(on=True):
    """Score tables of later sweeps may overwrite those of earlier ones (pinned staging buffers are recycled
    per size): for pipelines that consume a step's tables before the next step.  Off by default."""
    _runtime["reuse_result_buffers"] = bool(on)


class TorchCudaBackend:
    """Device buffers as torch CUDA tensors; all work on torch's current stream."""

    def __init__(self, device=None):
        import torch
        if not torch.cuda.is_available():
            raise _lib.CseLibraryError("no CUDA device: this path has no CPU fallback")
        self.torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self._dev_index = self.device.index
        self._current_stream_id = torch._C._cuda_getCurrentStream
        self._stream_handles = {}

    def empty(self, shape, dtype):
        t = self.torch
        td = {np.float32: t.float32, np.float64: t.float64, np.uint8: t.uint8, np.int32: t.int32}[np.dtype(dtype).type]
        return t.empty(shape, dtype=td, device=self.device)

    def zeros(self, shape, dtype):
        return self.empty(shape, dtype).zero_()

    def from_host(self, arr, pinned=False):
        t = self.torch.from_numpy(np.ascontiguousarray(arr))
        if pinned:
            t = t.pin_memory()
        return t.to(self.device, non_blocking=True)

    def to_host(self, buf):
        return buf.cpu().numpy()

    def ptr(self, buf):
        return None if buf is None else ctypes.c_void_p(buf.data_ptr())

    def ptr_at(self, buf, byte_offset):
        return ctypes.c_void_p(buf.data_ptr() + int(byte_offset))

    def stream(self):
        """Raw handle of torch's current stream on this device (one call per launch: the stream object and its
        ``c_void_p`` are cached per stream id instead of being rebuilt by ``torch.cuda.current_stream``)."""
        sid = self._current_stream_id(self._dev_index)[0]
        h = self._stream_handles.get(sid)
        if h is None:
            h = self._stream_handles[sid] = ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)
        return h

    def synchronize(self):
        self.torch.cuda.current_stream(self.device).synchronize()

    def event(self):
        e = self.torch.cuda.Event(enable_timing=True)
        e.record(self.torch.cuda.current_stream(self.device))
        return e

    def view_bytes_as(self, buf, dtype, tag=None):
        return self.staged_to_host(buf, tag=tag).view(dtype)

    def staged_to_host(self, buf, tag=None):
        """Device tensor -> numpy bytes.  With ``reuse_result_buffers`` (set by throughput drivers that consume
        a step's tables before the next step) the copy lands in a process-wide pinned staging buffer keyed by
        (device, size, tag), which the next copy with the same key overwrites; otherwise in a fresh pageable
        array.  ``tag`` names the result (algorithm, kind) so that two equally sized results of ONE step never
        share a buffer."""
        flat = buf.reshape(-1).view(self.torch.uint8)
        if not _runtime.get("reuse_result_buffers"):
            return flat.cpu().numpy()
        key = (self.device.index, flat.numel(), tag)
        host = _PINNED_POOL.get(key)
        if host is None:
            host = self.torch.empty((flat.numel(),), dtype=self.torch.uint8, pin_memory=True)
            _PINNED_POOL[key] = host
        host.copy_(flat, non_blocking=True)
        self.torch.cuda.current_stream(self.device).synchronize()
        return host.numpy()

    def slice_rows(self, buf, start, stop):
        return buf[start:stop]

    # --- chunk export on a side stream (PESQ path: candidate waveforms leave the device while the next chunk runs)
    def export_begin(self, parts, slot):
        """Enqueue device->host copies of ``parts`` = [(tensor, byte offset, nbytes)] on the copy stream, ordered
        after the work already enqueued on the current stream; pinned staging buffers are recycled per ``slot``."""
        t = self.torch
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = t.cuda.Stream(device=self.device)
            self._export_pinned = {}
        main = t.cuda.current_stream(self.device)
        ready = t.cuda.Event()
        ready.record(main)
        self._copy_stream.wait_event(ready)
        hosts = []
        with t.cuda.stream(self._copy_stream):
            for k, (buf, off, nbytes) in enumerate(parts):
                key = (slot, k)
                host = self._export_pinned.get(key)
                if host is None or host.numel() < nbytes:
                    host = self._export_pinned[key] = t.empty((nbytes,), dtype=t.uint8, pin_memory=True)
                src = buf.reshape(-1).view(t.uint8)[off:off + nbytes]
                host[:nbytes].copy_(src, non_blocking=True)
                hosts.append((host, nbytes))
            done = t.cuda.Event()
            done.record(self._copy_stream)
        return (hosts, done)

    def export_wait(self, handle):
        """Host arrays (fresh copies, the pinned buffers are reused) of a finished export."""
        hosts, done = handle
        done.synchronize()
        return [host[:n].numpy().copy() for host, n in hosts]

    def export_fence(self, handle):
        """The current stream may not overwrite the exported buffers before the copy has finished."""
        self.torch.cuda.current_stream(self.device).wait_event(handle[1])


def _synchronize_all(be):
    """Whole-device synchronisation (every stream), where the backend has one."""
    t = getattr(be, "torch", None)
    if t is not None:
        t.cuda.synchronize(be.device)
    else:
        be.synchronize()


class SweepEngine:
    """One batch of ``U`` equal-length (clean, noisy) pairs resident on one device."""

    def __init__(self, clean, noisy, sr=SR, lib=None, backend=None, chunk_items=DEFAULT_CHUNK_ITEMS, prepare_scoring=True):
        if sr != SR:
            raise ValueError("the sweep runs at 16 kHz (the reference resamples every pair to 16 kHz first)")
        self.lib = lib if lib is not None else (_runtime["lib"] or _lib.load())
        self.be = backend if backend is not None else (
            _runtime["backend_factory"]() if _runtime["backend_factory"] else TorchCudaBackend())
        self.real = self.lib.real
        noisy = np.atleast_2d(np.asarray(noisy))
        self.has_clean = clean is not None
        clean = np.atleast_2d(np.asarray(clean)) if self.has_clean else np.zeros_like(noisy)
        if clean.shape != noisy.shape or clean.ndim != 2:
            raise ValueError("clean and noisy must both be [U, L]")
        self.U, self.L = clean.shape
        if self.has_clean and prepare_scoring and self.L > self.lib.max_score_length(sr):
            raise _lib.CseError(_lib.CSE_EUNSUPPORTED, f"utterances of {self.L} samples exceed the scoring kernels' limit of "
                                f"{self.lib.max_score_length(sr)} samples (about 38 s at 16 kHz); split the recording")
        self.chunk_items = max(1, min(int(chunk_items), MAX_WAV_WORKSPACE_BYTES // max(1, self.L * np.dtype(self.real).itemsize)))
        be, lib_ = self.be, self.lib
        tkey = (id(lib_), getattr(be, "device", "host").__str__())
        self.tables = _TABLES_CACHE.get(tkey)
        if self.tables is None:        # constant tables: uploaded once per (library, device), shared by every engine
            self.tables = be.empty((lib_.tables_bytes(),), np.uint8)
            lib_.tables_init(be.ptr(self.tables), be.stream())
            be.synchronize()
            _TABLES_CACHE[tkey] = self.tables
        self.clean_host = np.ascontiguousarray(clean, dtype=self.real)      # kept for the host-side PESQ pool
        self.clean = be.from_host(self.clean_host)
        self.noisy = be.from_host(np.ascontiguousarray(noisy, dtype=self.real))
        self.h2d_bytes = 2 * clean.size * np.dtype(self.real).itemsize
        self._nf = {}
        self._stft = {}
        self._pow = {}
        self._noise = {}
        self._gamma = {}
        self._ws = {}
        self._plans = {}
        self._exports = []
        self.launches = 0
        self.cache = None
        self.sr = sr
        if self.has_clean and prepare_scoring:
            self.prepare_scoring()

    def prepare_scoring(self):
        """(Re)build the clean-side scoring caches (alignment spectra, VAD list, band envelopes)."""
        be, lib_ = self.be, self.lib
        rec = lib_.clean_cache_bytes(self.L, self.sr)
        if self.cache is None:
            self.cache = be.zeros((self.U * rec,), np.uint8)
        nbytes = lib_.clean_workspace_bytes(self.U, self.L, self.sr)
        ws = self._workspace("clean", nbytes)
        lib_.prepare_clean(be.ptr(self.tables), be.ptr(self.clean), self.U, self.L, self.sr, be.ptr(self.cache),
                           be.ptr(ws), nbytes, be.stream())
        self.launches += 3

    def reset(self):
        """Forget every derived cache (STFTs, noise PSDs, clean-side scoring caches) but keep the
        waveforms resident: the next sweep redoes the whole job from the raw signals."""
        self.drop_caches()
        if self.has_clean:
            self.prepare_scoring()

    # ------------------------------------------------------------------ optional kernel timing
    def enable_timing(self, on=True):
        """Bracket every chunk's enhance / score launches with events on the launching stream."""
        self._timing = [] if on else None

    def _tick(self):
        if getattr(self, "_timing", None) is None or not hasattr(self.be, "event"):
            return None
        return self.be.event()

    def _record(self, tag, n_items, e0, e1):
        if e0 is not None:
            self._timing.append((tag, n_items, e0, e1))

    def timing_summary(self):
        """{tag: (items, milliseconds)} after a synchronize; events are torch CUDA events."""
        out = {}
        for tag, n, e0, e1 in (getattr(self, "_timing", None) or []):
            it, ms = out.get(tag, (0, 0.0))
            out[tag] = (it + n, ms + e0.elapsed_time(e1))
        return out

    # ------------------------------------------------------------------ caches
    def _workspace(self, name, nbytes):
        nbytes = max(int(nbytes), 64)
        cur = self._ws.get(name)
        if cur is None or cur[1] < nbytes:
            self._ws[name] = (self.be.empty((nbytes,), np.uint8), nbytes)
        return self._ws[name][0]

    def n_frames(self, n_fft, hop):
        nf = self._nf.get(hop)
        if nf is None:
            nf = self._nf[hop] = self.lib.num_frames(self.L, hop)
        return nf

    def stft(self, n_fft, hop):
        """Y [U][nf][nbp] (complex as interleaved reals) of the noisy signals, cached per shape."""
        key = (n_fft, hop)
        if key not in self._stft:
            nf, nbp = self.n_frames(n_fft, hop), self.lib.bins_padded(n_fft)
            Y = self.be.empty((self.U, nf, nbp, 2), self.real)
            self.lib.stft_psd(self.be.ptr(self.tables), self.be.ptr(self.noisy), None, self.U, self.L, n_fft, hop,
                              0.0, self.be.ptr(Y), None, self.be.stream())
            self.launches += 1
            self._stft[key] = Y
        return self._stft[key]

    def _power(self, n_fft, hop):
        """|Y|^2 [U][nf][nbp] of the noisy signals, cached per shape (every estimator key of a shape reads it)."""
        key = (n_fft, hop)
        if key not in self._pow:
            nf, nbp = self.n_frames(n_fft, hop), self.lib.bins_padded(n_fft)
            P = self.be.empty((self.U, nf, nbp), self.real)
            self.lib.stft_psd(self.be.ptr(self.tables), self.be.ptr(self.noisy), None, self.U, self.L, n_fft, hop, 0.0,
                              None, self.be.ptr(P), self.be.stream())
            self.launches += 1
            self._pow[key] = P
        return self._pow[key]

    def noise(self, key):
        """Noise PSD for key = (n_fft, hop, method, percentile|None, eps) -> (buffer, time_varying)."""
        if key in self._noise:
            return self._noise[key]
        n_fft, hop, method, pct, eps = key
        be, lib_ = self.be, self.lib
        nf, nbp = self.n_frames(n_fft, hop), lib_.bins_padded(n_fft)
        if nf < 5 or method == "percentile":
            # short signals: every method falls back to the 25th percentile (noise_estimation.py:194-195)
            P = self._power(n_fft, hop)
            N = be.zeros((self.U, nbp), self.real)
            nbytes = lib_.noise_workspace_bytes(self.U, nf, n_fft)
            ws = self._workspace("noise", nbytes)
            lib_.noise_percentile(be.ptr(P), self.U, nf, n_fft, 20.0 if pct is None else pct, eps, be.ptr(N),
                                  be.ptr(ws), nbytes, be.stream())
            self.launches += 3
            out = (N, False)
        elif method == "min_tracking":
            P = self._power(n_fft, hop)
            N = be.zeros((self.U, nf, nbp), self.real)
            lib_.noise_mintrack(be.ptr(P), self.U, nf, n_fft, eps, be.ptr(N), None, 0, be.stream())
            self.launches += 1
            out = (N, True)
        elif method == "true_noise":
            if not self.has_clean:
                raise ValueError("TrueNoiseEstimator requires clean_audio and noisy_audio")
            N = be.zeros((self.U, nf, nbp), self.real)
            lib_.stft_psd(be.ptr(self.tables), be.ptr(self.noisy), be.ptr(self.clean), self.U, self.L, n_fft, hop,
                          eps, None, be.ptr(N), be.stream())
            self.launches += 1
            out = (N, True)
        else:
            raise ValueError(f"Unbekannte Methode: {method}")
        self._noise[key] = out
        return out

    def gamma(self, key6):
        """A-posteriori SNR [U][nf][nbp] for key6 = noise key + (effective noise_mu | None): ``cse_gamma`` of the
        cached STFT and noise PSD, shared by every candidate of the group (Wiener / MMSE / Log-MMSE)."""
        if key6 not in self._gamma:
            key, mu = key6[:5], key6[5]
            n_fft, hop = key[0], key[1]
            Y = self.stft(n_fft, hop)
            N, tv = self.noise(key)
            be = self.be
            G = be.empty((self.U, self.n_frames(n_fft, hop), self.lib.bins_padded(n_fft)), self.real)
            self.lib.gamma(be.ptr(Y), be.ptr(N), int(tv), self.U, self.L, n_fft, hop, -1.0 if mu is None else float(mu),
                           float(key[4]), be.ptr(G), be.stream())
            self.launches += 1
            self._gamma[key6] = G
        return self._gamma[key6]

    def gamma_many(self, keys6):
        """The missing ones of ``keys6`` by one ``cse_gamma_groups`` launch per n_fft (a one-pair sweep needs ~36 of
        them; one by one their dependent frame chains are the longest part of its device time)."""
        missing = [k for k in dict.fromkeys(keys6) if k not in self._gamma]
        if len(missing) < 2:
            return
        be = self.be
        by_nfft = {}
        for k in missing:
            by_nfft.setdefault(k[0], []).append(k)
        for n_fft, ks in by_nfft.items():
            descs = (_lib.GammaGroup * len(ks))()
            for d, k6 in zip(descs, ks):
                key, mu = k6[:5], k6[5]
                Y = self.stft(n_fft, key[1])
                N, tv = self.noise(key)
                G = be.empty((self.U, self.n_frames(n_fft, key[1]), self.lib.bins_padded(n_fft)), self.real)
                d.Y, d.N, d.G = be.ptr(Y).value, be.ptr(N).value, be.ptr(G).value
                d.noise_tv, d.hop, d.noise_mu, d.eps = int(tv), key[1], -1.0 if mu is None else float(mu), float(key[4])
                self._gamma[k6] = G
            self.lib.gamma_groups(self.U, self.L, n_fft, ctypes.cast(descs, ctypes.c_void_p), len(ks), be.stream())
            self.launches += -(-len(ks) // 32)

    def drop_caches(self):
        self._stft.clear()
        self._pow.clear()
        self._noise.clear()
        self._gamma.clear()

    # ------------------------------------------------------------------ the sweep
    def _plan(self, alg, points):
        """Grouped / deduplicated launch plan of a grid.  The host-side part is shared by every engine of the
        process through a bounded LRU keyed by the grid's CONTENT (algorithm, digest of the points, which shapes
        are time-varying) - never by object identity, so a rebuilt or edited point list can neither leak nor hit
        a stale plan; device-side maps / parameter uploads live in the per-engine copy."""
        sig = tuple(sorted({(n_fft, hop, self.n_frames(n_fft, hop) >= 5) for n_fft, hop, _ in points_shapes(points)}))
        self._validate_shapes(points)
        split_mu = bool(_runtime.get("gamma")) and alg != _lib.ALG_SS
        key = (alg, points_digest(points), len(points), sig, split_mu)
        pl = self._plans.get(key)
        if pl is not None:
            return pl
        out = _PLAN_CACHE.get(key)
        if out is not None:
            _PLAN_CACHE.move_to_end(key)
        else:
            groups = plan(alg, points, self.n_frames, split_mu=split_mu)
            info, col = [], 0
            for gkey, g in groups.items():
                n_rows = len(g["rows"])
                member_idx = np.concatenate([np.asarray(m, dtype=np.int64) for m in g["members"]])
                row_idx = np.concatenate([np.full(len(m), r, dtype=np.int64) for r, m in enumerate(g["members"])])
                info.append({"key": gkey, "rows": g["rows"], "params_host": _lib.pack_params(g["rows"]), "n_rows": n_rows,
                             "col0": col, "member_idx": member_idx, "row_idx": row_idx, "members": g["members"]})
                col += n_rows
            out = {"groups": info, "unique": col, "n_points": len(points), "gamma": split_mu}
            out["shared"] = out                       # the cached object itself: carries the per-device constants
            _PLAN_CACHE[key] = out
            while len(_PLAN_CACHE) > _PLAN_CACHE_MAX:
                _, old = _PLAN_CACHE.popitem(last=False)
                if old.get("dev"):
                    _synchronize_all(self.be)          # its device constants may still be read by enqueued kernels
        pl = dict(out)
        self._plans[key] = pl
        return pl

    def _plan_constants(self, pl):
        """Device-side constants of a plan - the packed parameter rows of every group and the nominal-table maps of
        ``cse_expand_scores`` for this batch size - uploaded ONCE per (plan, device) and shared by every engine and
        stream of the process (a corpus of one-pair buckets would otherwise re-upload ~70 small arrays per pair).
        The first upload is followed by a synchronisation, so that streams other than the uploading one may use
        the buffers without an event."""
        shared = pl["shared"]
        dkey = getattr(self.be, "device", "host").__str__()
        dev = shared.setdefault("dev", {}).get(dkey)
        fresh = False
        if dev is None:
            dev = shared["dev"][dkey] = {"params": [self.be.from_host(g["params_host"]) for g in pl["groups"]], "maps": OrderedDict()}
            fresh = True
        if self.U not in dev["maps"]:
            base = np.zeros(pl["n_points"], dtype=np.int32)
            stride = np.zeros(pl["n_points"], dtype=np.int32)
            for g in pl["groups"]:
                base[g["member_idx"]] = self.U * g["col0"] + g["row_idx"]
                stride[g["member_idx"]] = g["n_rows"]
            dev["maps"][self.U] = (self.be.from_host(base), self.be.from_host(stride))
            fresh = True
            while len(dev["maps"]) > 8:
                _synchronize_all(self.be)              # a kernel of another stream may still read the evicted maps
                dev["maps"].popitem(last=False)
        else:
            dev["maps"].move_to_end(self.U)
        if fresh:
            self.be.synchronize()
        return dev

    def _validate_shapes(self, points):
        """The build's operating range (include/cse.h), checked once per grid with a message that names the
        offending point instead of failing in the middle of a sweep."""
        for k in points_shapes(points):
            n_fft, hop, method = k
            if n_fft not in (256, 512, 1024, 2048):
                raise _lib.CseError(_lib.CSE_EINVAL, f"n_fft {n_fft} not in {{256,512,1024,2048}}")
            if hop <= 0 or hop % 2 or hop > n_fft // 2:
                raise _lib.CseError(_lib.CSE_EINVAL, f"hop {hop} must be even and <= n_fft/2 = {n_fft // 2}")
            if self.L <= n_fft // 2:
                raise _lib.CseError(_lib.CSE_EINVAL, f"length {self.L} must exceed n_fft/2 = {n_fft // 2} (reflect padding)")
            nf = self.n_frames(n_fft, hop)
            limit = 4096 if (method == "min_tracking" and nf >= 5) else (8192 if method in ("percentile", "min_tracking") else None)
            if limit is not None and nf > limit:
                raise _lib.CseError(_lib.CSE_EUNSUPPORTED, f"{nf} frames (hop {hop}) exceed the {method} estimator's limit of {limit}")

    def sweep_device(self, alg_name, points, u_pad=None, chunk_sink=None):
        """Enqueue the whole sweep of one algorithm; returns (device table, plan).

        ``chunk_sink(group, item0, wav, records)``, if given, receives every chunk's raw candidate waveforms
        ([n][L], host) and score records (lag, flags, ...) - the feed of the host-side PESQ pool.  The copies run on
        a side stream from one of two alternating waveform buffers while the next chunk is computed.

        The device table is the NOMINAL score table [u_pad][n_points] of 16-byte records (u_pad >= U
        rows so that equally sized tables can be all-gathered across ranks; rows >= U are zero).
        Unique candidates are scored once into a scratch buffer and broadcast to their duplicates by
        ``cse_expand_scores``; nothing is synchronised or copied to the host here."""
        alg = ALGORITHM_IDS[alg_name] if isinstance(alg_name, str) else int(alg_name)
        pl = self._plan(alg, points)
        u_pad = self.U if u_pad is None else int(u_pad)
        be, lib_ = self.be, self.lib
        rec = self.lib.score_dtype.itemsize
        uniq = be.empty((max(1, self.U * pl["unique"] * rec),), np.uint8)
        dev = self._plan_constants(pl)
        keep = [dev]
        rb = np.dtype(self.real).itemsize
        # One pair (the reference's own call pattern: optimize_parameters per pair and algorithm, and every bucket of
        # a variable-length corpus): a group is a few dozen to a few hundred candidates, far below one wave of the
        # score kernels.  With U == 1 the groups' score blocks are consecutive in the unique table, so every group
        # enhances into its slice of ONE waveform buffer and the whole grid is aligned and scored by one launch each.
        fused = (self.U == 1 and chunk_sink is None and _runtime.get("fuse_single", True) and 0 < pl["unique"] <= self.chunk_items
                 and pl["unique"] * self.L * rb <= MAX_WAV_WORKSPACE_BYTES)
        if fused:
            wav_all = self._workspace("wav", pl["unique"] * self.L * rb)
            t_first = self._tick()
            inputs = []
            if pl.get("gamma"):
                self.gamma_many([g["key"] for g in pl["groups"]])
            for g in pl["groups"]:
                Y = self.stft(g["key"][0], g["key"][1])
                inputs.append((Y,) + ((self.gamma(g["key"]), 2) if pl.get("gamma") else self.noise(g["key"])))
            # ... and the groups of one kernel instantiation (n_fft, kind of noise input) go out as ONE launch
            buckets = {}
            for gi, g in enumerate(pl["groups"]):
                buckets.setdefault((g["key"][0], int(inputs[gi][2])), []).append(gi)
            for (n_fft, tv), members in buckets.items():
                descs = (_lib.EnhanceGroup * len(members))()
                for d, gi in zip(descs, members):
                    g = pl["groups"][gi]
                    d.Y, d.N = be.ptr(inputs[gi][0]).value, be.ptr(inputs[gi][1]).value
                    d.params = be.ptr(dev["params"][gi]).value
                    d.out = be.ptr_at(wav_all, g["col0"] * self.L * rb).value
                    d.hop, d.n_params = g["key"][1], g["n_rows"]
                lib_.enhance_groups(be.ptr(self.tables), alg, tv, 1, self.L, n_fft, ctypes.cast(descs, ctypes.c_void_p),
                                    len(members), be.stream())
                self.launches += -(-len(members) // 24)
        for gi, g in enumerate(() if fused else pl["groups"]):
            key = g["key"]
            n_fft, hop = key[0], key[1]
            Y = self.stft(n_fft, hop)
            if pl.get("gamma"):
                N, tv = self.gamma(key), 2
            else:
                N, tv = self.noise(key)
            n_rows = g["n_rows"]
            params = dev["params"][gi]
            sc_ptr = be.ptr_at(uniq, self.U * g["col0"] * rec)          # group block [U][n_rows]
            total = self.U * n_rows
            # near-equal chunks, as many as the total rounds to: a group of 1.04 chunks runs as ONE launch (a few
            # hundred candidates launched on their own would cost two waves of the whole GPU), never more than 1.5 chunks
            n_chunks = max(1, int(total / self.chunk_items + 0.5))
            chunk = -(-total // n_chunks)
            if chunk * self.L * rb > MAX_WAV_WORKSPACE_BYTES:
                chunk = min(self.chunk_items, total)
            wavs = [self._workspace("wav", chunk * self.L * rb)]
            if chunk_sink is not None:
                wavs.append(self._workspace("wav1", chunk * self.L * rb))
            nbytes = lib_.score_workspace_bytes(chunk, self.L, SR)
            ws = self._workspace("score", nbytes)
            for k, i0 in enumerate(range(0, total, chunk)):
                n = min(chunk, total - i0)
                wav = wavs[k % len(wavs)]
                if chunk_sink is not None:
                    self._export_retire(chunk_sink, keep=1)     # chunk k-2 used this buffer: its copy must be over
                t0 = self._tick()
                lib_.enhance_items(be.ptr(self.tables), alg, be.ptr(Y), be.ptr(N), int(tv), self.L, n_fft, hop,
                                   be.ptr(params), n_rows, i0, n, be.ptr(wav), be.stream())
                t1 = self._tick()
                sargs = (be.ptr(self.tables), be.ptr(wav), i0, n, n_rows, self.L, SR, be.ptr(self.clean),
                         be.ptr(self.cache), 1, sc_ptr, be.ptr(ws), nbytes, be.stream())
                lib_.align_items(*sargs)
                t2 = self._tick()
                lib_.stoi_items(*sargs)
                t3 = self._tick()
                self._record(("enhance", alg, n_fft, hop, key[2]), n, t0, t1)
                self._record(("align", alg, n_fft, hop, key[2]), n, t1, t2)
                self._record(("stoi", alg, n_fft, hop, key[2]), n, t2, t3)
                self.launches += 3
                if chunk_sink is not None:
                    parts = [(wav, 0, n * self.L * rb), (uniq, (self.U * g["col0"] + i0) * rec, n * rec)]
                    handle = be.export_begin(parts, slot=k % 2) if hasattr(be, "export_begin") else [
                        np.array(be.to_host(wav).reshape(-1).view(np.uint8)[:parts[0][2]]),
                        np.array(be.to_host(uniq).reshape(-1).view(np.uint8)[parts[1][1]:parts[1][1] + parts[1][2]])]
                    self._exports.append((handle, g, i0, n))
            if chunk_sink is not None:
                self._export_retire(chunk_sink, keep=0)
        if fused:
            n = pl["unique"]
            nbytes = lib_.score_workspace_bytes(n, self.L, SR)
            ws = self._workspace("score", nbytes)
            t1 = self._tick()
            sargs = (be.ptr(self.tables), be.ptr(wav_all), 0, n, n, self.L, SR, be.ptr(self.clean), be.ptr(self.cache), 1,
                     be.ptr(uniq), be.ptr(ws), nbytes, be.stream())
            lib_.align_items(*sargs)
            t2 = self._tick()
            lib_.stoi_items(*sargs)
            t3 = self._tick()
            self._record(("enhance", alg, 0, 0, "one pair"), n, t_first, t1)
            self._record(("align", alg, 0, 0, "one pair"), n, t1, t2)
            self._record(("stoi", alg, 0, 0, "one pair"), n, t2, t3)
            self.launches += 2
        table = (be.zeros if u_pad > self.U else be.empty)((u_pad * pl["n_points"] * rec,), np.uint8)
        dbase, dstride = dev["maps"][self.U]
        lib_.expand_scores(be.ptr(uniq), be.ptr(dbase), be.ptr(dstride), self.U, pl["n_points"], be.ptr(table), be.stream())
        self.launches += 1
        self._keepalive = (keep, uniq)  # must outlive the enqueued kernels
        self.last_unique = pl["unique"]
        return table, pl

    def _export_retire(self, sink, keep):
        """Hand finished chunk exports to the sink until at most ``keep`` are in flight."""
        be = self.be
        while len(self._exports) > keep:
            handle, g, i0, n = self._exports.pop(0)
            if hasattr(be, "export_wait"):
                be.export_fence(handle)
                wav_b, rec_b = be.export_wait(handle)
            else:
                wav_b, rec_b = handle
            sink(g, i0, wav_b.view(self.real).reshape(n, self.L), rec_b.view(self.lib.score_dtype))

    def table_to_host(self, raw_bytes, pl, u_rows):
        """[u_pad][n_points] record bytes -> structured array [u_rows, n_points]."""
        dt = self.lib.score_dtype
        return np.asarray(raw_bytes).reshape(-1).view(dt).reshape(-1, pl["n_points"])[:u_rows]

    def select_device(self, table, n_points, pesq=None):
        """Enqueue the reference's three-way sequential selection (``speech_enhancement_comparison.py:186-216``)
        over a nominal device table [>= U][n_points]; returns the device buffer of ``cse_winner_t`` [U][3]
        (stoi, pesq, balance).  ``pesq``: host array [U, n_points] of float64 (NaN = candidate skipped) or None
        (PESQ = 0.0 everywhere: only the ``stoi`` winner is meaningful)."""
        be, lib_ = self.be, self.lib
        win = be.empty((self.U * 3 * lib_.winner_dtype.itemsize,), np.uint8)
        pq = None
        if pesq is not None:
            pesq = np.ascontiguousarray(pesq, dtype=np.float64)
            if pesq.shape != (self.U, n_points):
                raise ValueError(f"pesq must be [U, n_points] = {(self.U, n_points)}, got {pesq.shape}")
            pq = be.from_host(pesq)
        lib_.select_best(be.ptr(table), be.ptr(pq), self.U, int(n_points), be.ptr(win), be.stream())
        self.launches += 1
        self._keepalive_sel = pq
        return win

    def winners_to_host(self, win, tag=None):
        """Device winners buffer -> structured array [U, 3] of ``_lib.WINNER_DTYPE``."""
        return np.asarray(self.be.view_bytes_as(win, np.uint8, **({"tag": tag} if tag is not None else {}))).view(
            self.lib.winner_dtype).reshape(-1, 3)[:self.U]

    def sweep(self, alg_name, points):
        """Scores of every grid point for every utterance.

        Returns a structured array [U, len(points)] (stoi, snr, lag, flags) in the reference's
        grid order.  Points that differ only in dead parameters are computed once and their
        score broadcast (the reference would produce bit-identical duplicates).  ``unique`` is
        recorded in ``self.last_unique``.
        """
        table, pl = self.sweep_device(alg_name, points)
        return self.table_to_host(self.be.view_bytes_as(table, np.uint8), pl, self.U)

    def sweep_ranges(self, alg_name, param_ranges):
        points = grid_points(param_ranges)
        return points, self.sweep(alg_name, points)

    # ------------------------------------------------------------------ single candidates
    def enhance(self, alg_name, points):
        """Raw enhanced waveforms [U, len(points), L] (host, library precision) - used for the
        winners' artefacts and by the per-call drop-in functions."""
        alg = ALGORITHM_IDS[alg_name] if isinstance(alg_name, str) else int(alg_name)
        groups = plan(alg, points, self.n_frames)
        out = np.zeros((self.U, len(points), self.L), dtype=self.real)
        be, lib_ = self.be, self.lib
        for key, g in groups.items():
            n_fft, hop = key[0], key[1]
            Y = self.stft(n_fft, hop)
            N, tv = self.noise(key)
            rows = g["rows"]
            params = be.from_host(_lib.pack_params(rows))
            wav = be.empty((self.U * len(rows), self.L), self.real)
            lib_.enhance(be.ptr(self.tables), alg, be.ptr(Y), be.ptr(N), int(tv), self.U, self.L, n_fft, hop,
                         be.ptr(params), len(rows), be.ptr(wav), be.stream())
            self.launches += 1
            host = be.to_host(wav).reshape(self.U, len(rows), self.L)
            for r, members in enumerate(g["members"]):
                out[:, members, :] = host[:, r:r + 1, :]
        return out

    def enhance_list(self, alg_name, points, wanted):
        """Raw enhanced waveforms of a SPARSE set of candidates: ``wanted`` = [(utterance, grid index)] ->
        {(utterance, grid index): waveform [L]}.  One ``cse_enhance_list`` launch per noise-PSD group that holds
        wanted points - the winners of a dataset sweep (<= 3 per utterance and algorithm) cost a fraction of a
        percent of the sweep that found them."""
        alg = ALGORITHM_IDS[alg_name] if isinstance(alg_name, str) else int(alg_name)
        pl = self._plan(alg, points)
        if "point_loc" not in pl["shared"]:
            grp = np.zeros(pl["n_points"], dtype=np.int32)
            row = np.zeros(pl["n_points"], dtype=np.int32)
            for gi, g in enumerate(pl["groups"]):
                grp[g["member_idx"]] = gi
                row[g["member_idx"]] = g["row_idx"]
            pl["shared"]["point_loc"] = (grp, row)
        grp, row = pl["shared"]["point_loc"]
        dev = self._plan_constants(pl)
        be, lib_ = self.be, self.lib
        by_group = {}
        for u, i in wanted:
            by_group.setdefault(int(grp[i]), []).append((int(u), int(i)))
        out = {}
        for gi, lst in by_group.items():
            g = pl["groups"][gi]
            key = g["key"]
            n_fft, hop = key[0], key[1]
            Y = self.stft(n_fft, hop)
            if pl.get("gamma"):
                N, tv = self.gamma(key), 2
            else:
                N, tv = self.noise(key)
            uniq = sorted({(u, int(row[i])) for u, i in lst})                 # identical device candidates once
            items = np.array([u * g["n_rows"] + r for u, r in uniq], dtype=np.int32)
            params = dev["params"][gi]
            dev_items = be.from_host(items)
            wav = be.empty((len(items), self.L), self.real)
            lib_.enhance_list(be.ptr(self.tables), alg, be.ptr(Y), be.ptr(N), int(tv), self.L, n_fft, hop, be.ptr(params),
                              g["n_rows"], be.ptr(dev_items), len(items), be.ptr(wav), be.stream())
            self.launches += 1
            host = be.to_host(wav)
            slot = {ur: k for k, ur in enumerate(uniq)}
            for u, i in lst:
                out[(u, i)] = host[slot[(u, int(row[i]))]]
        return out

    def score_waveforms(self, wav, finalize=True):
        """Scores arbitrary waveforms [U, C, L] against this batch's clean signals."""
        wav = np.ascontiguousarray(wav, dtype=self.real)
        U, C, L = wav.shape
        if U != self.U or L != self.L:
            raise ValueError("waveforms must be [U, C, L] for this batch")
        return self._score_device(self.be.from_host(wav), C, finalize)

    def _score_device(self, dev, C, finalize):
        return self.scores_to_host(self._score_enqueue(dev, C, finalize), C)

    def scores_to_host(self, scores, C):
        return self.be.view_bytes_as(scores, self.lib.score_dtype).reshape(self.U, C)

    def _score_enqueue(self, dev, C, finalize):
        be, lib_ = self.be, self.lib
        U, L = self.U, self.L
        scores = be.empty((U * C * self.lib.score_dtype.itemsize,), np.uint8)
        nbytes = lib_.score_workspace_bytes(U * C, L, SR)
        ws = self._workspace("score", nbytes)
        lib_.score(be.ptr(self.tables), be.ptr(dev), U, C, L, SR, be.ptr(self.clean), be.ptr(self.cache),
                   int(finalize), be.ptr(scores), be.ptr(ws), nbytes, be.stream())
        self.launches += 2
        return scores

    def baseline_device(self):
        """:meth:`baseline` without the read-back: the device buffer, for :meth:`scores_to_host` (…, 1)[:, 0]."""
        return self._score_enqueue(self.noisy, 1, False)

    def baseline(self):
        """STOI / SNR of the unprocessed noisy signals (``optimize_parameters`` ``:116-118``), scored where they
        already are: on the device."""
        return self._score_device(self.noisy, 1, False)[:, 0]

    def noise_psd_host(self, method, n_fft, hop, percentile, eps):
        """(bins, 1) or (bins, frames) float64 per utterance, reference orientation."""
        from .grid import noise_key
        key = noise_key({"n_fft": n_fft, "hop_length": hop, "noise_method": method,
                         "noise_percentile": percentile}, eps)
        N, tv = self.noise(key)
        host = self.be.to_host(N).astype(np.float64)
        nb = n_fft // 2 + 1
        if tv:
            return [host[u, :, :nb].T.copy() for u in range(self.U)]
        return [host[u, :nb, None].copy() for u in range(self.U)]


def finalize_host(enhanced, lag, length):
    """``finalize_enhanced`` for a winner whose lag the device already estimated: shift, match
    length, clip (``speech_enhancement_comparison.py:62-67,29-36,105``)."""
    x = np.asarray(enhanced, dtype=np.float64)
    if lag > 0:
        x = np.pad(x, (lag, 0))
    elif lag < 0:
        x = x[-lag:]
    if len(x) > length:
        x = x[:length]
    elif len(x) < length:
        x = np.pad(x, (0, length - len(x)))
    return np.clip(x, -1.0, 1.0)
