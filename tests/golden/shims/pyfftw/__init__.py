"""Stub of pyfftw -> scipy.fft (test infrastructure only)."""


class config:
    NUM_THREADS = 1


def import_wisdom(w):
    return (True, True, True)


def export_wisdom():
    return (b"", b"", b"")


from . import interfaces  # noqa: E402,F401
