"""CPU restatement of the index algebra of the specialised back-transform kernels (csrc/fft24.cuh, fft48.cuh):
N = 3P = (3*R0)*Q, sparse radix-(3*R0) first stage of the R0 non-zero input blocks (+ the lone input f = +P/2),
in-place DIF stages over the 3*R0 sequences, digit-reversed read-out.  Checked against numpy's FFT of the
zero-padded input; the GPU tests compare the kernels themselves with the generic passes and the oracle."""
import numpy as np
import pytest

W = lambda n, m: np.exp(-2j * np.pi * m / n)


def sparse_first_stage(x_of_f, Q, R0):
    """buf[k2][n1] = w_N^{n1 k2} * sum_{n2} x[n1 + Q n2] w_{3R0}^{n2 k2},  k2 = r + 3q."""
    NS, N = 3 * R0, 3 * R0 * Q
    buf = np.zeros((NS, Q), complex)
    for r in range(3):
        for n1 in range(Q):
            n2 = [u if u < R0 // 2 else u - R0 for u in range(R0)]
            v = np.array([x_of_f(n1 + Q * m) * W(NS, m * r) for m in n2])
            o = np.fft.fft(v)                                   # position u = n2 mod R0
            if n1 == 0:                                         # f = +P/2: w_{3R0}^{(R0/2) k2} = w_6^r (-1)^q
                e = x_of_f((R0 // 2) * Q) * W(NS, (R0 // 2) * r)
                o = o + e * (-1.0) ** np.arange(R0)
            for q in range(R0):
                k2 = r + 3 * q
                buf[k2, n1] = o[q] * W(N, n1 * k2)
    return buf


def inplace_dif(buf, Q, stages):
    """In-place DIF stages; returns X with X[NS*k1 + k2] read from position rev(k1)."""
    NS = buf.shape[0]
    M = Q
    for R in stages[:-1]:
        sub = M // R
        for k2 in range(NS):
            for b in range(Q // M):
                for s in range(sub):
                    idx = [b * M + s + u * sub for u in range(R)]
                    o = np.fft.fft(buf[k2, idx])
                    for d in range(R):
                        buf[k2, idx[d]] = o[d] * W(M, s * d)
        M = sub
    RL, r0 = stages[-1], stages[0]
    NB = Q // RL
    X = np.zeros(NS * Q, complex)
    for k2 in range(NS):
        for b in range(NB):
            o = np.fft.fft(buf[k2, b * RL:(b + 1) * RL])
            k1lo = (b // (NB // r0)) + r0 * (b % (NB // r0)) if len(stages) == 3 else b
            for c in range(RL):
                X[NS * (k1lo + NB * c) + k2] = o[c]
    return X


@pytest.mark.parametrize("R0,LQ,stages", [
    (8, 5, (8, 4)), (8, 6, (8, 8)), (8, 7, (16, 8)), (8, 8, (16, 16)), (8, 9, (8, 8, 8)),   # fft24.cuh plans
    (16, 4, (16,)), (16, 5, (32,)),                                                         # fft48.cuh plans
])
def test_sparse_radix_decomposition(R0, LQ, stages):
    Q = 1 << LQ
    P, N = R0 * Q, 3 * R0 * Q
    rng = np.random.default_rng(LQ + R0)
    x = np.zeros(N, complex)
    vals = rng.normal(size=P + 1) + 1j * rng.normal(size=P + 1)
    for i, f in enumerate(range(-P // 2, P // 2 + 1)):         # the P+1 non-zero inputs |f| <= P/2
        x[f % N] = vals[i]
    ref = np.fft.fft(x)
    buf = sparse_first_stage(lambda f: x[f % N], Q, R0)
    got = inplace_dif(buf, Q, stages)
    assert np.abs(got - ref).max() <= 1e-13 * np.abs(ref).max()


def _pass_x_reads(nlx, nly, hs):
    """Index algebra of fft24_fetch_x (bldfm_b200/csrc/fft_herm.cuh): the elements S[row][col] that pass X of the
    real-output back-transform reads for H[fy][f] = (S[fy][f] + conj(S[-fy][-f]))/2, fy = 0..nly/2, |f| <= nlx/2;
    `hs`: the spectrum is conjugate-symmetric (half-plane march), interior elements are taken as H = S."""
    hx, hy, px, py = nlx // 2, nly // 2, (nlx - 1) // 2, (nly - 1) // 2
    reads = set()
    for fy in range(0, hy + 1):
        for f in range(-hx, hx + 1):
            one = hs and -hx < f < hx and 0 < fy < hy
            ok1 = fy <= py and -hx <= f <= px
            ok2 = (not one) and fy <= hy and -hx <= -f <= px
            if ok1:
                reads.add((fy, f if f >= 0 else f + nlx))
            if ok2:
                reads.add(((nly - fy) if fy > 0 else 0, (nlx - f) if f > 0 else -f))
    return reads


def test_pass_x_never_reads_the_mirror_stores_of_an_even_conjugate_symmetric_spectrum():
    """Why the half-plane march may skip its conjugate mirror stores (MarchArgs.skip_mirror): for even nlx, nly the
    pass-X read set of a conjugate-symmetric spectrum lies in the rows ky <= nly/2 plus the Nyquist column of the
    other rows (which the march writes as modes of their own, not as mirrors).  For odd sizes, or for a spectrum
    that is not known to be symmetric, mirror-row elements are read -- there the stores stay."""
    def mirror_reads(nlx, nly, hs):
        return {(r, c) for (r, c) in _pass_x_reads(nlx, nly, hs) if r > nly // 2 and not (nlx % 2 == 0 and c == nlx // 2)}
    for nlx, nly in ((8, 8), (16, 8), (8, 12), (64, 32)):
        assert not mirror_reads(nlx, nly, True)
        assert mirror_reads(nlx, nly, False)                       # general spectrum: both rows are combined
        # every marched element is still read (nothing else was dropped): rows 0..nly/2, all columns
        got = _pass_x_reads(nlx, nly, True)
        assert {(r, c) for r in range(nly // 2 + 1) for c in range(nlx)} <= got
    for nlx, nly in ((9, 8), (8, 9), (9, 9)):
        assert mirror_reads(nlx, nly, True)                        # odd sizes need mirror elements
