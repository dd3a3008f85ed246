"""ctypes binding of ``libdstr_b200.so`` (C-ABI in ``include/dstr_b200.h``).

There is NO CPU fallback: if the CUDA library is missing or no GPU is visible the
constructors raise.  The library is built in-tree by ``__graft_entry__.build()`` (or
``make -C aind_smartspim_destripe_b200/csrc``).
"""

from __future__ import annotations

import ctypes as C
import os
import threading
import weakref
from typing import Optional, Tuple

import numpy as np

_LIB_PATH = os.environ.get(
    "DSTR_LIBRARY", os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libdstr_b200.so")
)

DSTR_U16, DSTR_F32 = 0, 1
MODE_LOGSPACE, MODE_DISPATCH = 0, 1
FLAG_SHADOW, FLAG_EXPM1, FLAG_STACK_OTSU, FLAG_NO_SYNC = 1, 2, 4, 8
STAGE_NONE, STAGE_ANALYSIS, STAGE_OTSU, STAGE_FILTER, STAGE_SYNTH = 0, 1, 2, 3, 4
FETCH_CA, FETCH_CH, FETCH_STATS, FETCH_HIST = 0, 1, 2, 3
E_ARG, E_SHAPE, E_STATE, E_UNSUPPORTED = -1, -2, -3, -4
NUM_TIMERS = 9
TIMER_NAMES = (
    "analysis_l1",
    "analysis_deep",
    "histogram",
    "otsu",
    "row_filter",
    "synthesis_deep",
    "final_synthesis_epilogue",
    "chunk_total",
    "row_filter_level1",
)


class DstrParams(C.Structure):
    """``dstr_params``: the reference config dict {level, sigma, max_threshold}."""

    _fields_ = [("sigma", C.c_float), ("max_threshold", C.c_float), ("level", C.c_int)]


class EngineError(RuntimeError):
    pass


_lib = None
_lib_lock = threading.Lock()


def load_library() -> C.CDLL:
    """Load the CUDA library; raise loudly if it has not been built."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(_LIB_PATH):
            raise EngineError(
                f"{_LIB_PATH} not found: the CUDA extension is not built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'`). "
                "This engine has no CPU fallback."
            )
        lib = C.CDLL(_LIB_PATH)
        vp, ip, fp, dp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_double)
        pp = C.POINTER(DstrParams)
        sig = {
            "dstr_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]),
            "dstr_destroy": (C.c_int, [vp]),
            "dstr_last_error": (C.c_char_p, [vp]),
            "dstr_set_flat_dark": (C.c_int, [vp, vp, vp]),
            "dstr_filter_chunk": (
                C.c_int,
                [vp, vp, C.c_int, vp, C.c_int, C.c_int, pp, pp, C.c_float, C.c_int, C.c_int],
            ),
            "dstr_plane_stats": (C.c_int, [vp, vp, C.c_int, C.c_int, dp, dp, ip, C.c_float, C.c_float]),
            "dstr_flatfield_correction": (
                C.c_int,
                [C.c_int, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int],
            ),
            "dstr_max_level": (C.c_int, [C.c_int, C.c_int]),
            "dstr_level_shape": (C.c_int, [C.c_int, C.c_int, C.c_int, ip, ip]),
            "dstr_foreground_threshold": (C.c_float, [C.c_float]),
            "dstr_foreground_threshold_f32": (C.c_float, [C.c_float]),
            "dstr_notch_kernels": (C.c_int, [C.c_int, C.c_double, dp, dp]),
            "dstr_notch_design": (C.c_int, [C.c_int, C.c_double, C.c_double, ip]),
            "dstr_notch_apply_host": (C.c_int, [C.c_int, C.c_double, C.c_double, dp, dp]),
            "dstr_set_notch_tolerance": (C.c_int, [vp, C.c_double]),
            "dstr_host_alloc": (C.c_int, [C.POINTER(vp), C.c_uint64]),
            "dstr_host_free": (C.c_int, [vp]),
            "dstr_host_register": (C.c_int, [vp, C.c_uint64]),
            "dstr_host_unregister": (C.c_int, [vp]),
            "dstr_device_alloc": (C.c_int, [vp, C.POINTER(vp), C.c_uint64]),
            "dstr_device_free": (C.c_int, [vp, vp]),
            "dstr_memcpy_h2d": (C.c_int, [vp, vp, vp, C.c_uint64]),
            "dstr_memcpy_d2h": (C.c_int, [vp, vp, vp, C.c_uint64]),
            "dstr_synchronize": (C.c_int, [vp]),
            "dstr_compute_stream": (vp, [vp]),
            "dstr_set_profiling": (C.c_int, [vp, C.c_int]),
            "dstr_get_timers": (C.c_int, [vp, dp, C.POINTER(C.c_uint64)]),
            "dstr_reset_timers": (C.c_int, [vp]),
            "dstr_set_debug_stop": (C.c_int, [vp, C.c_int]),
            "dstr_debug_fetch": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_uint64]),
            "dstr_set_subchunk": (C.c_int, [vp, C.c_int]),
            "dstr_set_overlap": (C.c_int, [vp, C.c_int]),
            "dstr_set_tma": (C.c_int, [vp, C.c_int]),
            "dstr_set_umma": (C.c_int, [vp, C.c_int]),
            "dstr_set_row_filter": (C.c_int, [vp, C.c_int]),
            "dstr_notch_umma_info": (C.c_int, [C.c_int, C.c_double, ip]),
            "dstr_dual_band_chunk": (
                C.c_int,
                [vp, vp, C.c_int, vp, C.c_int, C.c_float, C.c_float, C.c_int, C.POINTER(C.c_float), C.c_float, C.c_float, vp],
            ),
            "dstr_histogram_u16": (C.c_int, [vp, vp, C.c_int, vp]),
            "dstr_png_unfilter": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp]),
            "dstr_notch_umma_apply_host": (C.c_int, [C.c_int, C.c_double, C.c_double, dp, dp, ip]),
            "dstr_downscale2x": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp]),
            "dstr_blosc_available": (C.c_int, [C.c_int]),
            "dstr_blosc_compress": (C.c_int64, [vp, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, vp, C.c_uint64]),
            "dstr_blosc_decompress": (C.c_int64, [vp, C.c_uint64, vp, C.c_uint64]),
            "dstr_set_pyramid_outputs": (C.c_int, [vp, vp, vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


EXPORTED_SYMBOLS = (
    "dstr_create dstr_destroy dstr_last_error dstr_set_flat_dark dstr_filter_chunk dstr_plane_stats "
    "dstr_flatfield_correction dstr_max_level dstr_level_shape dstr_foreground_threshold dstr_foreground_threshold_f32 "
    "dstr_notch_kernels dstr_notch_design dstr_notch_apply_host dstr_set_notch_tolerance dstr_host_alloc dstr_host_free dstr_host_register dstr_host_unregister "
    "dstr_device_alloc dstr_device_free dstr_memcpy_h2d dstr_memcpy_d2h dstr_synchronize "
    "dstr_compute_stream dstr_set_profiling dstr_get_timers dstr_reset_timers dstr_set_debug_stop "
    "dstr_debug_fetch dstr_set_subchunk dstr_set_overlap dstr_set_tma dstr_set_umma dstr_set_row_filter dstr_notch_umma_info "
    "dstr_notch_umma_apply_host dstr_downscale2x dstr_set_pyramid_outputs dstr_blosc_available dstr_blosc_compress "
    "dstr_blosc_decompress dstr_dual_band_chunk dstr_histogram_u16 dstr_png_unfilter"
).split()


def _raise(code: int, ctx=None, what: str = ""):
    lib = load_library()
    msg = lib.dstr_last_error(ctx)
    msg = msg.decode() if msg else ""
    text = f"{what}: error {code}: {msg}"
    if code == E_UNSUPPORTED:
        raise NotImplementedError(text)
    if code in (E_ARG, E_SHAPE, E_STATE):
        raise ValueError(text)
    raise EngineError(text)


def max_level(H: int, W: int) -> int:
    """pywt.dwtn_max_level((H, W), 'db3')."""
    return int(load_library().dstr_max_level(int(H), int(W)))


def level_shape(H: int, W: int, level: int) -> Tuple[int, int]:
    h, w = C.c_int(), C.c_int()
    rc = load_library().dstr_level_shape(int(H), int(W), int(level), C.byref(h), C.byref(w))
    if rc:
        _raise(rc, None, "dstr_level_shape")
    return h.value, w.value


def notch_kernels(n: int, s: float) -> Tuple[np.ndarray, np.ndarray]:
    hp = np.empty(n, dtype=np.float64)
    hq = np.empty(n, dtype=np.float64)
    dp = C.POINTER(C.c_double)
    rc = load_library().dstr_notch_kernels(int(n), float(s), hp.ctypes.data_as(dp), hq.ctypes.data_as(dp))
    if rc:
        _raise(rc, None, "dstr_notch_kernels")
    return hp, hq


def notch_design(n: int, s: float, eps: float = 1e-6) -> dict:
    info = (C.c_int * 6)()
    rc = load_library().dstr_notch_design(int(n), float(s), float(eps), info)
    if rc:
        _raise(rc, None, "dstr_notch_design")
    return dict(zip(("ntap_e", "ue_lo", "ntap_o", "uo_lo", "J", "Jpad"), [int(v) for v in info]))


def notch_apply_host(x: np.ndarray, s: float, eps: float = 1e-6) -> np.ndarray:
    """B x from the device's own float32 tables (host evaluation, double accumulation)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    dp = C.POINTER(C.c_double)
    rc = load_library().dstr_notch_apply_host(x.size, float(s), float(eps), x.ctypes.data_as(dp), y.ctypes.data_as(dp))
    if rc:
        _raise(rc, None, "dstr_notch_apply_host")
    return y


def notch_umma_info(n: int, s: float = 0.0) -> dict:
    """Geometry of the tensor-core row filter for a band of width ``n`` (tables sized for notch width ``s``)."""
    info = (C.c_int * 8)()
    rc = load_library().dstr_notch_umma_info(int(n), float(s), info)
    if rc:
        _raise(rc, None, "dstr_notch_umma_info")
    keys = ("eligible", "passes", "outputs_per_pass", "k_chunks", "table_bytes", "smem_bytes", "outputs", "k_padded")
    return dict(zip(keys, [int(v) for v in info]))


def notch_umma_apply_host(x: np.ndarray, s: float, thr: float):
    """B x through the data path of the tensor-core kernel, evaluated on the host (test support)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    dp = C.POINTER(C.c_double)
    info = (C.c_int * 4)()
    rc = load_library().dstr_notch_umma_apply_host(
        x.size, float(s), float(thr), x.ctypes.data_as(dp), y.ctypes.data_as(dp), info
    )
    if rc:
        _raise(rc, None, "dstr_notch_umma_apply_host")
    return y, {"Rb": int(info[0]), "r3": int(info[1]), "mmas_per_item": int(info[2]), "tb_compact": int(info[3])}


def foreground_threshold(threshold_mask: float = 0.3) -> float:
    return float(load_library().dstr_foreground_threshold(float(threshold_mask)))


def foreground_threshold_f32(threshold_mask: float = 0.3) -> float:
    """The float16 rule as a threshold on the float32 pixel value (NaN = never)."""
    return float(load_library().dstr_foreground_threshold_f32(float(threshold_mask)))


def make_params(cfg: Optional[dict]) -> Optional[DstrParams]:
    """Reference config dict -> ``dstr_params`` (run_capsule.py:377-388)."""
    if cfg is None:
        return None
    wavelet = cfg.get("wavelet", "db3")
    if wavelet != "db3":
        raise NotImplementedError(
            f"wavelet '{wavelet}' is not implemented by the B200 engine (db3 only; no CPU fallback)"
        )
    level = cfg.get("level", 0)
    sigma = cfg.get("sigma", 64)
    max_threshold = cfg.get("max_threshold", 4)
    if sigma <= 0:
        raise ValueError("sigma must be positive")  # filtering.py:111-112 via notch()
    if level is not None and level < 0:
        raise ValueError(f"Level value of {level} is too low . Minimum level is 0.")
    return DstrParams(float(sigma), float(max_threshold), -1 if level is None else int(level))


class PinnedBuffer:
    """Page-locked host buffer exposed as a numpy array (dstr_host_alloc)."""

    def __init__(self, shape, dtype):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self._ptr = C.c_void_p()
        rc = load_library().dstr_host_alloc(C.byref(self._ptr), max(nbytes, 1))
        if rc:
            _raise(rc, None, "dstr_host_alloc")
        buf = (C.c_ubyte * max(nbytes, 1)).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self._ptr is not None and self._ptr.value:
            self.array = None
            load_library().dstr_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceBuffer:
    """Raw device allocation owned by an engine (benchmark: HBM-resident chunks)."""

    def __init__(self, engine: "DestripeEngine", nbytes: int):
        self.engine = engine
        self.nbytes = int(nbytes)
        self.ptr = C.c_void_p()
        rc = engine.lib.dstr_device_alloc(engine.ctx, C.byref(self.ptr), self.nbytes)
        if rc:
            _raise(rc, engine.ctx, "dstr_device_alloc")

    def upload(self, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        rc = self.engine.lib.dstr_memcpy_h2d(self.engine.ctx, self.ptr, arr.ctypes.data_as(C.c_void_p), arr.nbytes)
        if rc:
            _raise(rc, self.engine.ctx, "dstr_memcpy_h2d")

    def download(self, shape, dtype) -> np.ndarray:
        out = np.empty(shape, dtype=dtype)
        assert out.nbytes <= self.nbytes
        rc = self.engine.lib.dstr_memcpy_d2h(self.engine.ctx, out.ctypes.data_as(C.c_void_p), self.ptr, out.nbytes)
        if rc:
            _raise(rc, self.engine.ctx, "dstr_memcpy_d2h")
        return out

    def free(self):
        if self.ptr is not None and self.ptr.value:
            self.engine.lib.dstr_device_free(self.engine.ctx, self.ptr)
            self.ptr = None


def _np_dtype_code(dt: np.dtype) -> int:
    if dt == np.uint16:
        return DSTR_U16
    if dt == np.float32:
        return DSTR_F32
    raise ValueError(f"unsupported buffer dtype {dt}")


class DestripeEngine:
    """One GPU context for planes of a fixed (H, W); not thread-safe (one per thread)."""

    def __init__(self, H: int, W: int, max_planes: int = 16, device: int = 0):
        self.lib = load_library()
        self.H, self.W, self.max_planes, self.device = int(H), int(W), int(max_planes), int(device)
        self.ctx = C.c_void_p()
        rc = self.lib.dstr_create(self.device, self.max_planes, self.H, self.W, C.byref(self.ctx))
        if rc:
            self.ctx = None
            _raise(rc, None, "dstr_create")
        self.max_level = max_level(self.H, self.W)
        self._flat_dark_key = None

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "ctx", None):
            self.lib.dstr_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int, what: str):
        if rc:
            if not getattr(self, "ctx", None):
                raise EngineError(f"{what}: this engine has been closed (get_engine evicts and closes engines; "
                                  "do not keep one across calls of the functional API)")
            _raise(rc, self.ctx, what)

    # -- shadow correction fields --------------------------------------------------------------
    def set_flat_dark(self, flat: Optional[np.ndarray], dark: Optional[np.ndarray]):
        self._shadow_key = None  # filtering._engine_with_shadow's cache describes what IT uploaded last
        if flat is None or dark is None:
            self._ck(self.lib.dstr_set_flat_dark(self.ctx, None, None), "dstr_set_flat_dark")
            self._flat_dark_key = None
            return
        key = (id(flat), id(dark))
        if key == self._flat_dark_key:
            return
        f = np.ascontiguousarray(flat, dtype=np.float32)
        d = np.ascontiguousarray(dark, dtype=np.float32)
        if f.shape != (self.H, self.W) or d.shape != (self.H, self.W):
            raise ValueError(f"flat/dark must be ({self.H}, {self.W}); got {f.shape} / {d.shape}")
        self._ck(
            self.lib.dstr_set_flat_dark(self.ctx, f.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p)),
            "dstr_set_flat_dark",
        )
        self._flat_dark_key = key
        self._flat_dark_refs = (flat, dark)  # keep ids alive

    # -- hot path ------------------------------------------------------------------------------
    def filter_chunk_ptr(self, in_ptr, in_code, out_ptr, out_code, Z, cells, no_cells, high_int, mode, flags):
        pc = C.byref(cells) if cells is not None else None
        rc = self.lib.dstr_filter_chunk(
            self.ctx, in_ptr, in_code, out_ptr, out_code, int(Z), pc, C.byref(no_cells), float(high_int), mode, flags
        )
        self._ck(rc, "dstr_filter_chunk")

    def filter_chunk(
        self,
        data: np.ndarray,
        no_cells: DstrParams,
        cells: Optional[DstrParams] = None,
        out: Optional[np.ndarray] = None,
        out_dtype=np.uint16,
        high_int: float = 2700.0,
        mode: int = MODE_LOGSPACE,
        flags: int = 0,
    ) -> np.ndarray:
        """Filter a (Z, H, W) host array (uint16 or float32)."""
        if data.ndim != 3 or data.shape[1:] != (self.H, self.W):
            raise ValueError(f"expected (Z, {self.H}, {self.W}) planes, got {data.shape}")
        data = np.ascontiguousarray(data)
        in_code = _np_dtype_code(data.dtype)
        if out is None:
            out = np.empty(data.shape, dtype=out_dtype)
        if not out.flags.c_contiguous or out.shape != data.shape:
            raise ValueError("out must be C-contiguous with the input's shape")
        out_code = _np_dtype_code(out.dtype)
        self.filter_chunk_ptr(
            data.ctypes.data_as(C.c_void_p),
            in_code,
            out.ctypes.data_as(C.c_void_p),
            out_code,
            data.shape[0],
            cells,
            no_cells,
            high_int,
            mode,
            flags,
        )
        return out

    def dual_band_chunk(self, data: np.ndarray, sigma_fg: float, sigma_bg: float, thresholds, level: int = -1,
                        crossover: float = 10.0, dark: float = 0.0, flat: Optional[np.ndarray] = None,
                        out: Optional[np.ndarray] = None) -> np.ndarray:
        """Classic dual-band filter of a (Z, H, W) host chunk (``dstr_dual_band_chunk``) -> uint16."""
        if data.ndim != 3 or data.shape[1:] != (self.H, self.W):
            raise ValueError(f"expected (Z, {self.H}, {self.W}) planes, got {data.shape}")
        data = np.ascontiguousarray(data)
        Z = data.shape[0]
        thr = np.ascontiguousarray(np.broadcast_to(np.asarray(thresholds, dtype=np.float32), (Z,)))
        if out is None:
            out = np.empty(data.shape, dtype=np.uint16)
        if out.dtype != np.uint16 or not out.flags.c_contiguous or out.shape != data.shape:
            raise ValueError("out must be a C-contiguous uint16 array with the input's shape")
        fptr = None
        if flat is not None:
            flat = np.ascontiguousarray(flat, dtype=np.float32)
            if flat.shape != (self.H, self.W):
                raise ValueError(f"flat must have shape ({self.H}, {self.W})")
            fptr = flat.ctypes.data_as(C.c_void_p)
        rc = self.lib.dstr_dual_band_chunk(
            self.ctx, data.ctypes.data_as(C.c_void_p), _np_dtype_code(data.dtype), out.ctypes.data_as(C.c_void_p), Z,
            float(sigma_fg), float(sigma_bg), int(level), thr.ctypes.data_as(C.POINTER(C.c_float)), float(crossover),
            float(dark), fptr,
        )
        self._ck(rc, "dstr_dual_band_chunk")
        return out

    def dual_band_chunk_ptr(self, in_ptr, in_code, out_ptr, Z, sigma_fg, sigma_bg, thresholds, level=-1, crossover=10.0,
                            dark=0.0, flat: Optional[np.ndarray] = None):
        """``dstr_dual_band_chunk`` on raw (host or device) pointers; ``thresholds``: (Z,) float32 host array."""
        thr = np.ascontiguousarray(thresholds, dtype=np.float32)
        if thr.shape != (int(Z),):
            raise ValueError("one threshold per plane")
        fptr = None
        if flat is not None:
            flat = np.ascontiguousarray(flat, dtype=np.float32)
            fptr = flat.ctypes.data_as(C.c_void_p)
        rc = self.lib.dstr_dual_band_chunk(self.ctx, in_ptr, in_code, out_ptr, int(Z), float(sigma_fg), float(sigma_bg),
                                           int(level), thr.ctypes.data_as(C.POINTER(C.c_float)), float(crossover),
                                           float(dark), fptr)
        self._ck(rc, "dstr_dual_band_chunk")

    def histogram_u16(self, data: np.ndarray) -> np.ndarray:
        """Exact per-plane histogram of a (Z, H, W) uint16 chunk -> (Z, 65536) uint32."""
        if data.ndim != 3 or data.shape[1:] != (self.H, self.W) or data.dtype != np.uint16:
            raise ValueError(f"expected (Z, {self.H}, {self.W}) uint16 planes")
        data = np.ascontiguousarray(data)
        hist = np.empty((data.shape[0], 65536), dtype=np.uint32)
        rc = self.lib.dstr_histogram_u16(self.ctx, data.ctypes.data_as(C.c_void_p), data.shape[0],
                                         hist.ctypes.data_as(C.c_void_p))
        self._ck(rc, "dstr_histogram_u16")
        return hist

    def plane_stats(self, data: np.ndarray, high_int: float = 2700.0, threshold_mask: float = 0.3):
        if data.ndim != 3 or data.shape[1:] != (self.H, self.W):
            raise ValueError(f"expected (Z, {self.H}, {self.W}) planes, got {data.shape}")
        data = np.ascontiguousarray(data)
        Z = data.shape[0]
        fg = np.zeros(Z, dtype=np.float64)
        bg = np.zeros(Z, dtype=np.float64)
        uc = np.zeros(Z, dtype=np.int32)
        dp = C.POINTER(C.c_double)
        rc = self.lib.dstr_plane_stats(
            self.ctx,
            data.ctypes.data_as(C.c_void_p),
            _np_dtype_code(data.dtype),
            Z,
            fg.ctypes.data_as(dp),
            bg.ctypes.data_as(dp),
            uc.ctypes.data_as(C.POINTER(C.c_int)),
            float(high_int),
            float(threshold_mask),
        )
        self._ck(rc, "dstr_plane_stats")
        return fg, bg, uc

    # -- instrumentation ---------------------------------------------------------------------------
    def set_profiling(self, enabled: bool):
        self._ck(self.lib.dstr_set_profiling(self.ctx, 1 if enabled else 0), "dstr_set_profiling")

    def reset_timers(self):
        self._ck(self.lib.dstr_reset_timers(self.ctx), "dstr_reset_timers")

    def timers(self):
        ms = (C.c_double * NUM_TIMERS)()
        n = C.c_uint64()
        self._ck(self.lib.dstr_get_timers(self.ctx, ms, C.byref(n)), "dstr_get_timers")
        return {k: float(ms[i]) for i, k in enumerate(TIMER_NAMES)}, int(n.value)

    def synchronize(self):
        self._ck(self.lib.dstr_synchronize(self.ctx), "dstr_synchronize")

    def compute_stream(self) -> int:
        return int(self.lib.dstr_compute_stream(self.ctx) or 0)

    def set_subchunk(self, planes: int):
        self._ck(self.lib.dstr_set_subchunk(self.ctx, int(planes)), "dstr_set_subchunk")

    def set_notch_tolerance(self, eps: float):
        self._ck(self.lib.dstr_set_notch_tolerance(self.ctx, float(eps)), "dstr_set_notch_tolerance")

    def downscale2x(self, vol: np.ndarray) -> np.ndarray:
        """uint16 (Z, H, W) -> (Z//2, H//2, W//2): truncated 2x2x2 windowed mean (one pyramid level)."""
        vol = np.ascontiguousarray(vol)
        if vol.dtype != np.uint16 or vol.ndim != 3:
            raise ValueError("downscale2x expects a uint16 (Z, H, W) array")
        Z, H, W = vol.shape
        out = np.empty((Z // 2, H // 2, W // 2), dtype=np.uint16)
        self._ck(self.lib.dstr_downscale2x(self.ctx, vol.ctypes.data_as(C.c_void_p), Z, H, W,
                                           out.ctypes.data_as(C.c_void_p)), "dstr_downscale2x")
        return out

    def set_pyramid_outputs(self, level1: Optional[np.ndarray], level2: Optional[np.ndarray] = None):
        """Host arrays that receive pyramid levels 1 / 2 of the next filtered uint16 chunks."""
        p1 = level1.ctypes.data_as(C.c_void_p) if level1 is not None else None
        p2 = level2.ctypes.data_as(C.c_void_p) if level2 is not None else None
        self._ck(self.lib.dstr_set_pyramid_outputs(self.ctx, p1, p2), "dstr_set_pyramid_outputs")
        self._pyr_refs = (level1, level2)

    def set_tma(self, enabled: bool):
        self._ck(self.lib.dstr_set_tma(self.ctx, 1 if enabled else 0), "dstr_set_tma")

    def set_umma(self, enabled: bool):
        """Row filter on the tcgen05 / TMEM / TMA kernel (opt-in: parity-green but measured slower than the default
        ``mma.sync`` kernel)."""
        self._ck(self.lib.dstr_set_umma(self.ctx, 1 if enabled else 0), "dstr_set_umma")

    def set_row_filter(self, kind: int):
        """0: register-tiled FMA row filter, 1 (default): mma.sync row filter (8 rows per block), 2: the same with 4 rows
        per block (the fallback form)."""
        self._ck(self.lib.dstr_set_row_filter(self.ctx, int(kind)), "dstr_set_row_filter")

    def set_overlap(self, enabled: bool):
        self._ck(self.lib.dstr_set_overlap(self.ctx, 1 if enabled else 0), "dstr_set_overlap")

    def set_debug_stop(self, stage: int):
        self._ck(self.lib.dstr_set_debug_stop(self.ctx, int(stage)), "dstr_set_debug_stop")

    def debug_fetch(self, what: int, level: int, Z: int) -> np.ndarray:
        h, w = level_shape(self.H, self.W, level)
        if what in (FETCH_CA, FETCH_CH):
            out = np.empty((Z, h, w), dtype=np.float32)
        elif what == FETCH_STATS:
            out = np.empty((Z, 8), dtype=np.float32)
        elif what == FETCH_HIST:
            out = np.empty((Z, 256), dtype=np.uint32)
        else:
            raise ValueError("unknown fetch item")
        self._ck(
            self.lib.dstr_debug_fetch(self.ctx, what, level, out.ctypes.data_as(C.c_void_p), out.nbytes),
            "dstr_debug_fetch",
        )
        return out


class _EngineCache(dict):
    """dict that can be weakly referenced (and hashed by identity): the per-thread cache dies with its thread, and
    its engines with it"""

    __hash__ = object.__hash__
    __eq__ = object.__eq__


_tls = threading.local()  # per-thread engine cache (an engine is not thread-safe)
_all_caches = weakref.WeakSet()  # for release_engines(); weak, so caches of finished threads are collected
_engines_lock = threading.Lock()
_MAX_CACHED_ENGINES = 6   # each engine owns a device workspace proportional to max_planes * H * W


def default_device() -> int:
    """GPU index for the functional API: DSTR_DEVICE, else LOCAL_RANK (one process per GPU), else 0."""
    return int(os.environ.get("DSTR_DEVICE", os.environ.get("LOCAL_RANK", "0")))


def get_engine(H: int, W: int, device: Optional[int] = None, max_planes: int = 16) -> DestripeEngine:
    """Per-thread cached engine for a plane shape (the functional API uses this).

    The cache is a small per-thread LRU: the reference API is stateless, so callers that sweep
    many plane shapes must not accumulate one device workspace per shape, and an eviction can only
    close an engine of the calling thread (which is not inside a call at that moment).
    """
    if device is None:
        device = default_device()
    cache = getattr(_tls, "engines", None)
    if cache is None:
        cache = _tls.engines = _EngineCache()
        with _engines_lock:
            _all_caches.add(cache)
    key = (int(device), int(H), int(W))
    eng = cache.pop(key, None)
    if eng is not None and eng.max_planes < max_planes:
        eng.close()
        eng = None
    if eng is None:
        while len(cache) >= _MAX_CACHED_ENGINES:
            cache.pop(next(iter(cache))).close()  # least recently used
        eng = DestripeEngine(H, W, max_planes=max_planes, device=device)
    cache[key] = eng  # most recently used last
    return eng


def release_engines():
    """Close every cached engine (call only when no thread is inside the functional API)."""
    with _engines_lock:
        for cache in list(_all_caches):
            for eng in list(cache.values()):
                eng.close()
            cache.clear()
