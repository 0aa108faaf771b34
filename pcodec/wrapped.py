"""pcodec.wrapped (pco_python/src/wrapped/): FileCompressor, ChunkCompressor, FileDecompressor, ChunkDecompressor."""
from pcodec_b200.wrapped import *  # noqa: F401,F403
from pcodec_b200.wrapped import ChunkCompressor, ChunkDecompressor, FileCompressor, FileDecompressor  # noqa: F401
