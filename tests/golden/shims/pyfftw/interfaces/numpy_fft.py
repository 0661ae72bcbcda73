import scipy.fft as _sf


def fft2(a, s=None, axes=(-2, -1), norm="backward", **kw):
    return _sf.fft2(a, s=s, axes=axes, norm=norm)


def ifft2(a, s=None, axes=(-2, -1), norm="backward", **kw):
    return _sf.ifft2(a, s=s, axes=axes, norm=norm)
