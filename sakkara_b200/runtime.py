"""
ctypes binding of ``libskk_b200.so`` (C ABI in ``include/skk_b200.h``).

There is no CPU fallback: if the shared library has not been built, or no CUDA device is present,
creating a plan raises.  Build with ``python -m sakkara_b200.build_ext`` (or ``__graft_entry__.build()``).
"""
import ctypes as C
import os
from typing import List, Optional, Tuple

import numpy as np

from sakkara_b200.plan.kernel_plan import KernelPlan

ABI_VERSION = 1
EVAL_DEVICE_IO = 1
EVAL_TIME_KERNELS = 2
EVAL_NO_SYNC = 4

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'lib', 'libskk_b200.so')


class TermDesc(C.Structure):
    _fields_ = [('level', C.c_int32), ('xcol', C.c_int32), ('n_slots', C.c_int64), ('theta_index', C.c_void_p)]


class BlockDesc(C.Structure):
    _fields_ = [('family', C.c_int32), ('reserved', C.c_int32), ('offset', C.c_int64), ('size', C.c_int64),
                ('a_const', C.c_void_p), ('a_const_len', C.c_int64), ('a_index', C.c_void_p),
                ('b_const', C.c_void_p), ('b_const_len', C.c_int64), ('b_index', C.c_void_p)]


class PlanDesc(C.Structure):
    _fields_ = [('abi_version', C.c_int32), ('dtype', C.c_int32), ('family', C.c_int32), ('device', C.c_int32),
                ('n_rows', C.c_int64), ('n_params', C.c_int64), ('max_batch', C.c_int32), ('n_xcols', C.c_int32),
                ('xcols', C.POINTER(C.c_void_p)), ('y', C.c_void_p), ('y_kind', C.c_int32),
                ('sec_code_bytes', C.c_int32), ('eta_offset', C.c_void_p), ('n_primary', C.c_int64),
                ('seg_offsets', C.c_void_p), ('n_secondary', C.c_int64), ('sec_codes', C.c_void_p),
                ('n_terms', C.c_int32), ('n_blocks', C.c_int32), ('terms', C.POINTER(TermDesc)),
                ('blocks', C.POINTER(BlockDesc)), ('sigma_const', C.c_double), ('sigma_theta_index', C.c_int64),
                ('logp_const', C.c_double), ('n_rows_total', C.c_int64), ('include_priors', C.c_int32),
                ('reserved', C.c_int32)]


class Stats(C.Structure):
    _fields_ = [('evals', C.c_uint64), ('launches', C.c_uint64), ('row_kernel_launches', C.c_uint64),
                ('last_row_kernel_ms', C.c_double), ('last_eval_ms', C.c_double),
                ('stored_bytes_per_row', C.c_uint64), ('device_bytes', C.c_uint64), ('grid', C.c_int32),
                ('block', C.c_int32), ('smem_bytes', C.c_int32), ('n_tiles', C.c_int32),
                ('rows_per_tile', C.c_int32), ('kernel_variant', C.c_int32)]

    def as_dict(self) -> dict:
        return {name: getattr(self, name) for name, _ in self._fields_}


EXPORTS = ('skk_plan_create', 'skk_eval', 'skk_plan_destroy', 'skk_last_error', 'skk_comm_unique_id',
           'skk_comm_init', 'skk_comm_destroy', 'skk_peer_alloc', 'skk_peer_init', 'skk_stats', 'skk_stream',
           'skk_trace_enable', 'skk_trace_read', 'skk_alu_peak', 'skk_abi_version', 'skk_leapfrog')

_lib = None


def load_library(path: Optional[str] = None) -> C.CDLL:
    """Load the shared library (once per process) and declare the prototypes."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or os.environ.get('SKK_B200_LIB', LIB_PATH)
    if not os.path.exists(path):
        raise RuntimeError(f'{path} not found: the CUDA extension is not built '
                           f'(run `python -m sakkara_b200.build_ext`); there is no CPU fallback')
    lib = C.CDLL(path)
    lib.skk_plan_create.argtypes = [C.POINTER(PlanDesc), C.POINTER(C.c_void_p)]
    lib.skk_plan_create.restype = C.c_int
    lib.skk_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_uint32]
    lib.skk_eval.restype = C.c_int
    lib.skk_plan_destroy.argtypes = [C.c_void_p]
    lib.skk_plan_destroy.restype = C.c_int
    lib.skk_last_error.argtypes = []
    lib.skk_last_error.restype = C.c_char_p
    lib.skk_comm_unique_id.argtypes = [C.c_void_p]
    lib.skk_comm_unique_id.restype = C.c_int
    lib.skk_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
    lib.skk_comm_init.restype = C.c_int
    lib.skk_comm_destroy.argtypes = [C.c_void_p]
    lib.skk_comm_destroy.restype = C.c_int
    lib.skk_peer_alloc.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
    lib.skk_peer_alloc.restype = C.c_int
    lib.skk_peer_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
    lib.skk_peer_init.restype = C.c_int
    lib.skk_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    lib.skk_stats.restype = C.c_int
    lib.skk_stream.argtypes = [C.c_void_p]
    lib.skk_stream.restype = C.c_void_p
    lib.skk_leapfrog.argtypes = [C.c_void_p] * 6 + [C.c_int32, C.c_double, C.c_double]
    lib.skk_leapfrog.restype = C.c_int
    lib.skk_trace_enable.argtypes = [C.c_void_p, C.c_int32]
    lib.skk_trace_enable.restype = C.c_int
    lib.skk_trace_read.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    lib.skk_trace_read.restype = C.c_int64
    lib.skk_alu_peak.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_double)]
    lib.skk_alu_peak.restype = C.c_int
    lib.skk_abi_version.argtypes = []
    lib.skk_abi_version.restype = C.c_int
    if lib.skk_abi_version() != ABI_VERSION:
        raise RuntimeError(f'libskk_b200.so has ABI {lib.skk_abi_version()}, binding expects {ABI_VERSION}')
    _lib = lib
    return lib


def alu_peak(device: int = 0) -> dict:
    """Measured thread-instruction rates of the device: fp64 FMA, fp32 FMA, MUFU.EX2 (per second)."""
    lib = load_library()
    out = {}
    for kind, name in enumerate(('dfma_per_s', 'ffma_per_s', 'mufu_per_s')):
        v = C.c_double()
        _check(lib, lib.skk_alu_peak(device, kind, C.byref(v)))
        out[name] = v.value
    return out


class SkkError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f'libskk_b200 error {code}: {message}')
        self.code = code


def _check(lib, code: int) -> None:
    if code != 0:
        raise SkkError(code, (lib.skk_last_error() or b'').decode(errors='replace'))


def _ptr(a: Optional[np.ndarray]) -> Optional[int]:
    return None if a is None else a.ctypes.data


class SkkPlan:
    """A device-resident plan.  Bound to the creating process; not fork-safe (create lazily)."""

    def __init__(self, kp: KernelPlan, max_batch: int = 1, device: int = 0):
        self.lib = load_library()
        self.kernel_plan_info = kp.describe()
        self.dtype = kp.dtype
        self.n_params = kp.n_params
        self.max_batch = int(max_batch)
        self.n_rows = kp.n_rows
        self.stored_bytes_per_row = kp.stored_bytes_per_row
        self._pid = os.getpid()
        self._handle = C.c_void_p()

        keep: List[np.ndarray] = []             # host arrays must outlive skk_plan_create only

        def hold(a, dt=None):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=dt)
            keep.append(a)
            return a

        xcols = [hold(x, kp.dtype) for x in kp.xcols]
        xptrs = (C.c_void_p * max(1, len(xcols)))(*[x.ctypes.data for x in xcols])
        terms = (TermDesc * max(1, len(kp.terms)))()
        for i, t in enumerate(kp.terms):
            ti = hold(t.theta_index, np.int32)
            terms[i] = TermDesc(t.level, t.xcol, len(ti), ti.ctypes.data)
        blocks = (BlockDesc * max(1, len(kp.blocks)))()
        for i, b in enumerate(kp.blocks):
            ac, ai = hold(b.a_const, np.float64), hold(b.a_index, np.int32)
            bc, bi = hold(b.b_const, np.float64), hold(b.b_index, np.int32)
            blocks[i] = BlockDesc(b.family, 0, b.offset, b.size, _ptr(ac), 0 if ac is None else len(ac), _ptr(ai),
                                  _ptr(bc), 0 if bc is None else len(bc), _ptr(bi))
        y = hold(kp.y)
        off = hold(kp.eta_offset, kp.dtype)
        seg = hold(kp.seg_offsets, np.int64)
        codes = hold(kp.sec_codes)
        desc = PlanDesc(
            abi_version=ABI_VERSION, dtype=1 if kp.dtype == np.float64 else 0, family=kp.family, device=int(device),
            n_rows=kp.n_rows, n_params=kp.n_params, max_batch=self.max_batch, n_xcols=len(xcols), xcols=xptrs,
            y=_ptr(y), y_kind=kp.y_kind, sec_code_bytes=0 if codes is None else codes.dtype.itemsize,
            eta_offset=_ptr(off), n_primary=kp.n_primary, seg_offsets=_ptr(seg) if kp.n_primary else None,
            n_secondary=kp.n_secondary, sec_codes=_ptr(codes), n_terms=len(kp.terms), n_blocks=len(kp.blocks),
            terms=terms, blocks=blocks, sigma_const=kp.sigma_const, sigma_theta_index=kp.sigma_theta_index,
            logp_const=kp.logp_const, n_rows_total=kp.n_rows_total, include_priors=1 if kp.include_priors else 0,
            reserved=0)
        _check(self.lib, self.lib.skk_plan_create(C.byref(desc), C.byref(self._handle)))
        del keep

    # -- evaluation ----------------------------------------------------------------------------
    def eval(self, theta: np.ndarray, time_kernels: bool = False) -> Tuple[np.ndarray, np.ndarray]:
        """``theta [B, P]`` (or ``[P]``) host array -> ``(logp [B], dlogp [B, P])`` host arrays."""
        self._guard()
        theta = np.ascontiguousarray(theta, dtype=self.dtype)
        single = theta.ndim == 1
        th = theta.reshape(1, -1) if single else theta
        if th.ndim != 2 or th.shape[1] != self.n_params:
            raise ValueError(f'theta must have shape [B, {self.n_params}]')
        batch = th.shape[0]
        logp = np.empty(batch, dtype=self.dtype)
        grad = np.empty((batch, self.n_params), dtype=self.dtype)
        flags = EVAL_TIME_KERNELS if time_kernels else 0
        _check(self.lib, self.lib.skk_eval(self._handle, th.ctypes.data, batch, logp.ctypes.data, grad.ctypes.data, flags))
        return (logp[0], grad[0]) if single else (logp, grad)

    def eval_device(self, theta_ptr: int, batch: int, logp_ptr: int, grad_ptr: int, sync: bool = True,
                    time_kernels: bool = False) -> None:
        """Evaluate with device-resident theta and outputs (raw device pointers, e.g. torch ``data_ptr()``)."""
        self._guard()
        flags = EVAL_DEVICE_IO | (0 if sync else EVAL_NO_SYNC) | (EVAL_TIME_KERNELS if time_kernels else 0)
        _check(self.lib, self.lib.skk_eval(self._handle, C.c_void_p(theta_ptr), batch, C.c_void_p(logp_ptr),
                                           C.c_void_p(grad_ptr), flags))

    def leapfrog(self, q_ptr: int, r_ptr: int, grad_ptr: int, eps_ptr: int, inv_mass_ptr: int, chains: int,
                 a: float, b: float) -> None:
        """``r += a*eps*grad; q += b*eps*inv_mass*r`` for ``chains`` device-resident chains, enqueued on the plan's stream."""
        self._guard()
        _check(self.lib, self.lib.skk_leapfrog(self._handle, C.c_void_p(q_ptr), C.c_void_p(r_ptr), C.c_void_p(grad_ptr),
                                               C.c_void_p(eps_ptr), C.c_void_p(inv_mass_ptr), chains, a, b))

    # -- multi-GPU -------------------------------------------------------------------------------
    def unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        _check(self.lib, self.lib.skk_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, nranks: int) -> None:
        self._guard()
        if len(unique_id) != 128:
            raise ValueError('unique_id must be 128 bytes')
        buf = C.create_string_buffer(unique_id, 128)
        _check(self.lib, self.lib.skk_comm_init(self._handle, buf, rank, nranks))

    def peer_alloc(self, nranks: int) -> bytes:
        """Allocate the receive buffer of the peer-memory all-reduce; returns its 64-byte IPC handle."""
        self._guard()
        buf = C.create_string_buffer(64)
        _check(self.lib, self.lib.skk_peer_alloc(self._handle, nranks, buf))
        return buf.raw

    def peer_init(self, handles, rank: int, nranks: int) -> None:
        """``handles``: the IPC handles of all ranks in rank order (from an all-gather)."""
        self._guard()
        blob = b''.join(handles)
        if len(blob) != 64 * nranks:
            raise ValueError('need one 64-byte handle per rank')
        _check(self.lib, self.lib.skk_peer_init(self._handle, C.create_string_buffer(blob, len(blob)), rank, nranks))

    def comm_destroy(self) -> None:
        self._guard()
        _check(self.lib, self.lib.skk_comm_destroy(self._handle))

    # -- misc -------------------------------------------------------------------------------------
    def stats(self) -> dict:
        self._guard()
        st = Stats()
        _check(self.lib, self.lib.skk_stats(self._handle, C.byref(st)))
        return st.as_dict()

    @property
    def stream(self) -> int:
        self._guard()
        return int(self.lib.skk_stream(self._handle) or 0)

    def trace_enable(self, on: bool = True) -> None:
        """Per-warp timeline of the row kernel (diagnostics; see ``skk_trace_enable`` in the header)."""
        self._guard()
        _check(self.lib, self.lib.skk_trace_enable(self._handle, 1 if on else 0))

    def trace_read(self, finish: bool = False) -> np.ndarray:
        """
        ``[grid * warps, 16]`` uint64 of the row kernel: globaltimer ns at entry / prologue done / first tile / loop
        end / exit (words 5-7: tiles per class; 8-11: value tables filled / priors done / all warps of the CTA done
        / global partials stored).  ``finish=True``: the records of the kernel behind it instead,
        one per CTA that ran -- entry / row kernel's results visible / own sums stored / exit / CTA class.
        """
        self._guard()
        st = self.stats()
        n_rows = st['grid'] * (st['block'] // 32)
        out = np.zeros((n_rows + 1024, 16), dtype=np.uint64)
        n = self.lib.skk_trace_read(self._handle, out.ctypes.data, out.size)
        if n < 0:
            _check(self.lib, int(n))
        if finish:
            tail = out[n_rows:]
            return tail[tail[:, 0] > 0]
        return out[:n_rows]

    def _guard(self) -> None:
        if not self._handle:
            raise RuntimeError('plan has been destroyed')
        if os.getpid() != self._pid:
            raise RuntimeError('a plan cannot be used in a forked process; create it in the worker')

    def close(self) -> None:
        if getattr(self, '_handle', None) and self._handle.value and os.getpid() == self._pid:
            self.lib.skk_plan_destroy(self._handle)
        self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
