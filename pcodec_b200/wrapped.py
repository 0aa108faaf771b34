"""Mirror of pcodec.wrapped (pco_python/src/wrapped/{compressor,decompressor}.rs) over the C-ABI of libcpcodec.so.

One page per chunk this round: `FileCompressor.chunk_compressor` raises PcoError("Unsupported") when the config's
PagingSpec would cut the chunk into several pages."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import ChunkConfig, Progress


def _lib_wrapped():
    L = _lib.lib()
    if not getattr(L, "_wrapped_ready", False):
        for name in ("pco_b200_chunk_compressor_n_pages", "pco_b200_chunk_compressor_meta_size"):
            getattr(L, name).restype = C.c_size_t
            getattr(L, name).argtypes = [C.c_void_p]
        for name in ("pco_b200_chunk_compressor_page_n", "pco_b200_chunk_compressor_page_size"):
            getattr(L, name).restype = C.c_size_t
            getattr(L, name).argtypes = [C.c_void_p, C.c_size_t]
        L.pco_b200_chunk_compressor_free.argtypes = [C.c_void_p]
        L.pco_b200_chunk_compressor_free.restype = None
        L._wrapped_ready = True
    return L


class ChunkCompressor:
    """pcodec.wrapped.ChunkCompressor (pco_python/src/wrapped/compressor.rs:22-115)."""

    def __init__(self, handle):
        self._h = handle

    def __del__(self):
        if getattr(self, "_h", None):
            _lib_wrapped().pco_b200_chunk_compressor_free(self._h)
            self._h = None

    def n_per_page(self):
        L = _lib_wrapped()
        return [L.pco_b200_chunk_compressor_page_n(self._h, i) for i in range(L.pco_b200_chunk_compressor_n_pages(self._h))]

    def write_meta(self):
        L = _lib_wrapped()
        buf = np.empty(L.pco_b200_chunk_compressor_meta_size(self._h), dtype=np.uint8)
        n = C.c_size_t()
        _lib.check(L.pco_b200_chunk_compressor_write_meta(self._h, buf.ctypes.data_as(C.c_void_p), C.c_size_t(buf.size), C.byref(n)))
        return buf[: n.value].tobytes()

    def write_page(self, page_idx):
        L = _lib_wrapped()
        buf = np.empty(max(L.pco_b200_chunk_compressor_page_size(self._h, page_idx), 1), dtype=np.uint8)
        n = C.c_size_t()
        _lib.check(L.pco_b200_chunk_compressor_write_page(self._h, C.c_size_t(page_idx), buf.ctypes.data_as(C.c_void_p), C.c_size_t(buf.size), C.byref(n)))
        return buf[: n.value].tobytes()


class FileCompressor:
    """pcodec.wrapped.FileCompressor (pco_python/src/wrapped/compressor.rs:15-91)."""

    def write_header(self):
        buf = np.empty(2, dtype=np.uint8)
        n = C.c_size_t()
        _lib.check(_lib_wrapped().pco_b200_file_compressor_write_header(buf.ctypes.data_as(C.c_void_p), C.c_size_t(2), C.byref(n)))
        return buf[: n.value].tobytes()

    def chunk_compressor(self, nums, config=None):
        arr = np.ascontiguousarray(nums)
        cfg = (config or ChunkConfig())._to_c()
        h = C.c_void_p()
        _lib.check(_lib_wrapped().pco_b200_chunk_compressor_new(arr.ctypes.data_as(C.c_void_p), C.c_size_t(arr.size), C.c_ubyte(_lib.dtype_byte(arr.dtype)),
                                                                 C.byref(cfg), C.byref(h)))
        return ChunkCompressor(h)


class ChunkDecompressor:
    """pcodec.wrapped.ChunkDecompressor (pco_python/src/wrapped/decompressor.rs:25-122): holds the chunk meta bytes."""

    def __init__(self, meta, dtype):
        self._meta, self._dtype = bytes(meta), np.dtype(dtype)

    def read_page_into(self, src, page_n, dst):
        """Decodes one page of `page_n` numbers from the head of `src` into `dst`; returns (Progress, bytes_read)."""
        if np.dtype(dst.dtype) != self._dtype:
            raise TypeError("dst dtype does not match the chunk's")
        sb = np.frombuffer(bytes(src), dtype=np.uint8)
        mb = np.frombuffer(self._meta, dtype=np.uint8)
        prog = _lib._CProgress()
        nread = C.c_size_t()
        _lib.check(_lib_wrapped().pco_b200_page_decompress(mb.ctypes.data_as(C.c_void_p), C.c_size_t(mb.size), sb.ctypes.data_as(C.c_void_p), C.c_size_t(sb.size),
                                                            C.c_size_t(page_n), C.c_ubyte(_lib.dtype_byte(self._dtype)), dst.ctypes.data_as(C.c_void_p),
                                                            C.c_size_t(dst.size), C.byref(prog), C.byref(nread)))
        return Progress(prog.n_processed, bool(prog.finished)), nread.value


class FileDecompressor:
    """pcodec.wrapped.FileDecompressor (pco_python/src/wrapped/decompressor.rs:16-95)."""

    @staticmethod
    def new(src):
        """Returns (FileDecompressor, bytes_read) from the wrapped header at the head of `src`."""
        sb = np.frombuffer(bytes(src), dtype=np.uint8)
        n = C.c_size_t()
        _lib.check(_lib_wrapped().pco_b200_file_decompressor_read_header(sb.ctypes.data_as(C.c_void_p), C.c_size_t(sb.size), C.byref(n)))
        return FileDecompressor(), n.value

    def chunk_decompressor(self, src, dtype):
        """Returns (ChunkDecompressor, bytes_read) from the chunk meta at the head of `src`."""
        sb = np.frombuffer(bytes(src), dtype=np.uint8)
        n = C.c_size_t()
        _lib.check(_lib_wrapped().pco_b200_chunk_meta_size(sb.ctypes.data_as(C.c_void_p), C.c_size_t(sb.size), C.c_ubyte(_lib.dtype_byte(np.dtype(dtype))), C.byref(n)))
        return ChunkDecompressor(sb[: n.value].tobytes(), dtype), n.value
