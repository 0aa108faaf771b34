"""pcodec.standalone (pco_python/src/standalone.rs:44-135): simple_compress, simple_decompress, simple_decompress_into, FileCompressor ..."""
from pcodec_b200.standalone import *  # noqa: F401,F403
from pcodec_b200.standalone import simple_compress, simple_decompress, simple_decompress_into  # noqa: F401
