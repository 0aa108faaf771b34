"""Mirror of `pcodec.standalone` (pco_python/src/standalone.rs:44-135) over libcpcodec.so."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import ChunkConfig, PcoError, Progress  # noqa: F401


def _src_buf(data):
    mv = memoryview(data)
    n = mv.nbytes
    arr = np.frombuffer(mv, dtype=np.uint8) if n else np.zeros(1, dtype=np.uint8)
    return arr, n


def simple_decompress_into(src, dst):
    """pcodec.standalone.simple_decompress_into (pco_python/src/standalone.rs:90-118): decompresses into a
    numpy array, returning Progress; never errors on a too-short or too-long dst."""
    return decompress_into_with_index(src, dst, None)


def simple_decompress(src, dtype=None, index=None):
    """pcodec.standalone.simple_decompress (pco_python/src/standalone.rs:120-135).  The reference infers the
    dtype from the file; pass `dtype` to assert it.  `index` is an optional side index (bytes)."""
    if dtype is None:
        dtype = peek_dtype(src)
        if dtype is None:
            return None
    # size the destination from n_hint when present, growing if the file holds more.  n_hint is untrusted (a 20-byte file may claim 2^64
    # numbers): the first allocation is capped at 2^32 bytes like the reference's (pco/src/standalone/constants.rs:10-17, decompressor.rs:265-269)
    cap = max(min(n_hint(src), (1 << 32) // np.dtype(dtype).itemsize), 1)
    while True:
        dst = np.empty(cap, dtype=dtype)
        prog = decompress_into_with_index(src, dst, index)
        if prog.finished:
            return dst[: prog.n_processed]
        cap = max(cap * 2, 1 << 16)


def decompress_into_with_index(src, dst, index=None):
    L = _lib.lib()
    arr, n = _src_buf(src)
    prog = _lib._CProgress()
    if index is not None:
        iarr, ilen = _src_buf(index)
        iptr, ilen_c = iarr.ctypes.data_as(C.c_void_p), C.c_size_t(ilen)
    else:
        iptr, ilen_c = None, C.c_size_t(0)
    rc = L.pco_b200_decompress_ex(arr.ctypes.data_as(C.c_void_p), C.c_size_t(n), C.c_ubyte(_lib.dtype_byte(dst.dtype)),
                                  dst.ctypes.data_as(C.c_void_p), C.c_size_t(dst.size), C.byref(prog), iptr, ilen_c,
                                  C.c_uint32(0), None)
    _lib.check(rc)
    return Progress(prog.n_processed, bool(prog.finished))


def build_index(src, dtype, n_total_hint=None):
    """Builds the per-batch side index of a standalone file on the device (one serial tANS walk)."""
    L = _lib.lib()
    arr, n = _src_buf(src)
    hint = n_total_hint if n_total_hint is not None else max(n_hint(src), 1)
    cap = L.pco_b200_index_size_bound(hint, max(hint >> 8, 64))
    while True:
        buf = np.empty(cap, dtype=np.uint8)
        used = C.c_size_t()
        rc = L.pco_b200_build_index(arr.ctypes.data_as(C.c_void_p), C.c_size_t(n), C.c_ubyte(_lib.dtype_byte(dtype)),
                                    buf.ctypes.data_as(C.c_void_p), C.c_size_t(cap), C.byref(used), C.c_uint32(0), None)
        if rc == 4:  # Io: index buffer too small
            cap *= 4
            continue
        _lib.check(rc)
        return buf[: used.value].tobytes()


def _header(src):
    """(standalone_version, uniform_type, n_hint, first_chunk_byte) — docs/format.md:173-192."""
    b = bytes(memoryview(src)[:32])
    if len(b) < 5 or b[:4] != b"pco!":
        return None
    ver = b[4]
    if ver < 2:
        return ver, 0, 0, 4 + 1
    pos = 5
    uniform = 0
    if ver >= 3:
        uniform = b[pos] if pos < len(b) else 0
        pos += 1
    wide = int.from_bytes(b[pos:pos + 10].ljust(10, b"\0"), "little")
    power = 1 + (wide & 63)
    n = (wide >> 6) & ((1 << power) - 1)
    pos += (6 + power + 7) // 8
    major = b[pos] if pos < len(b) else 0
    pos += 2 if major >= 4 else 1
    return ver, uniform, n, pos


def n_hint(src):
    h = _header(src)
    return h[2] if h else 0


def peek_dtype(src):
    h = _header(src)
    if h is None:
        raise PcoError("Corruption", "magic header does not match")
    if h[1]:
        return _lib.BYTE_TO_NP[h[1]]
    mv = memoryview(src)
    if h[3] >= mv.nbytes:
        raise PcoError("InsufficientData", "unable to peek number type from empty bytes")
    t = mv[h[3]]
    if t == 0:
        return None
    if t not in _lib.BYTE_TO_NP:
        raise PcoError("Corruption", f"peeked unknown number type byte: {t}")
    return _lib.BYTE_TO_NP[t]


def _compress(nums, config, uniform_type, want_index):
    L = _lib.lib()
    arr = np.ascontiguousarray(nums)
    dt = _lib.dtype_byte(arr.dtype)
    cfg = (config or ChunkConfig())._to_c()
    cap = L.pco_standalone_guarantee_file_size(arr.size, dt)
    if cfg.paging_spec == 1 or (cfg.max_page_n and cfg.max_page_n < (1 << 18)):
        # the guarantee above assumes default paging; small pages add a per-chunk meta allowance
        n_pages = cfg.n_exact_pages if cfg.paging_spec == 1 else -(-arr.size // max(cfg.max_page_n, 1))
        cap += int(n_pages) * 160
    dst = np.empty(cap, dtype=np.uint8)
    n_written = C.c_size_t()
    if want_index:
        icap = L.pco_b200_index_size_bound(arr.size, max(arr.size >> 8, 64))
        if cfg.paging_spec == 1 or cfg.max_page_n:
            icap += 64 * (int(cfg.n_exact_pages) if cfg.paging_spec == 1 else -(-arr.size // max(cfg.max_page_n, 1)))
        idx = np.empty(icap, dtype=np.uint8)
        ilen = C.c_size_t()
        iptr, icap_c, ilen_p = idx.ctypes.data_as(C.c_void_p), C.c_size_t(icap), C.byref(ilen)
    else:
        idx, ilen, iptr, icap_c, ilen_p = None, None, None, C.c_size_t(0), None
    rc = L.pco_b200_compress_ex(arr.ctypes.data_as(C.c_void_p), C.c_size_t(arr.size), C.c_ubyte(dt), C.byref(cfg),
                                C.c_int(1 if uniform_type else 0), dst.ctypes.data_as(C.c_void_p), C.c_size_t(cap), C.byref(n_written),
                                iptr, icap_c, ilen_p, C.c_uint32(0), None)
    _lib.check(rc)
    out = dst[: n_written.value].tobytes()
    if want_index:
        return out, idx[: ilen.value].tobytes()
    return out


def simple_compress(nums, config=None):
    """pcodec.standalone.simple_compress (pco_python/src/standalone.rs:44-88; pco/src/standalone/simple.rs:58-91)."""
    return _compress(nums, config, False, False)


def simple_compress_into(nums, config=None):
    """pco::standalone::simple_compress_into flavour (uniform-type header, pco/src/standalone/simple.rs:22-48)."""
    return _compress(nums, config, True, False)


def simple_compress_with_index(nums, config=None, uniform_type=False):
    """simple_compress plus the per-batch side index the GPU decoder uses (metadata beside the .pco bytes)."""
    return _compress(nums, config, uniform_type, True)


def decompress_chunks(src, dtype, chunk_offsets, chunk_ns):
    """Batched, index-free decompress of chunks at known byte offsets (pco_b200_decompress_chunks): `src` is a standalone
    file or bare chunks back to back; chunk_offsets[i] = byte offset of chunk i's type byte, chunk_ns[i] = its count.
    The device walks every chunk's tANS stream in parallel to build the per-batch index, then decodes."""
    L = _lib.lib()
    arr, n = _src_buf(src)
    offs = np.ascontiguousarray(chunk_offsets, dtype=np.uint64)
    ns = np.ascontiguousarray(chunk_ns, dtype=np.uint32)
    if offs.size != ns.size:
        raise ValueError("chunk_offsets and chunk_ns differ in length")
    dst = np.empty(int(ns.sum(dtype=np.uint64)), dtype=dtype)
    n_written = C.c_size_t()
    rc = L.pco_b200_decompress_chunks(arr.ctypes.data_as(C.c_void_p), C.c_size_t(n), C.c_ubyte(_lib.dtype_byte(dst.dtype)),
                                      offs.ctypes.data_as(C.c_void_p), ns.ctypes.data_as(C.c_void_p), C.c_size_t(offs.size),
                                      dst.ctypes.data_as(C.c_void_p), C.c_size_t(dst.size), C.byref(n_written), C.c_uint32(0), None)
    _lib.check(rc)
    return dst[: n_written.value]


class _CModeChoice(C.Structure):
    _fields_ = [("mode_spec", C.c_uint32), ("float_quant_k", C.c_uint32), ("float_mult_base", C.c_double), ("float_mult_inv_base", C.c_double),
                ("int_mult_base", C.c_uint64), ("bits_saved_per_num", C.c_double)]


def choose_mode(nums):
    """What `ModeSpec.auto()` resolves to in the reference for this chunk of numbers (pco/src/data_types/unsigned.rs:28-35,
    float.rs:70-98): a ModeSpec to put into a ChunkConfig.  Host-side planner logic of libcpcodec.so (pco_b200_choose_mode); needs no
    device.  `.inv_base` on a FloatMult answer is the inverse the reference's splitter multiplies by."""
    arr = np.ascontiguousarray(nums)
    out = _CModeChoice()
    _lib.check(_lib.lib().pco_b200_choose_mode(C.c_void_p(arr.ctypes.data), C.c_size_t(arr.size), C.c_ubyte(_lib.dtype_byte(arr.dtype)), C.byref(out)))
    spec = _lib.ModeSpec(out.mode_spec, base=out.float_mult_base, k=out.float_quant_k, int_base=out.int_mult_base)
    spec.inv_base = out.float_mult_inv_base
    spec.bits_saved_per_num = out.bits_saved_per_num
    return spec
