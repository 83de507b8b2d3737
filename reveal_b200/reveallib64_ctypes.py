"""`reveallib64` -- the reference's 64-bit build of the extension (setup.py:30-34,
-DSA64: saidx_t = int64, lcp_t = uint32, reveal.h:7-10).  Same numbers in wider
integers; the 32-bit total-length guard of addsequence (interface.c:61-68) is off."""
from .reveallib_ctypes import error, index as _index32


class index(_index32):
    _bits = 64


__all__ = ["index", "error"]
