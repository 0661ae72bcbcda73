from . import cache, numpy_fft  # noqa: F401
