"""NumPy model of the Stockham mixed-radix line FFT and the real<->half-complex packing used by
admp_b200/csrc/fft.cu (index math check; tests/test_fft_model.py compares it with numpy.fft)."""
import numpy as np

RADICES = (8, 4, 2, 3, 5, 7, 11, 13)


def factorize(n):
    """Radix sequence (largest primes first, then powers of two as 4s/8s)."""
    out = []
    for r in (13, 11, 7, 5, 3):
        while n % r == 0:
            out.append(r)
            n //= r
    while n % 8 == 0 and n > 8:
        out.append(8)
        n //= 8
    while n % 4 == 0:
        out.append(4)
        n //= 4
    while n % 2 == 0:
        out.append(2)
        n //= 2
    if n != 1:
        return None
    return out


def dft_small(x, sign):
    r = x.shape[0]
    k = np.arange(r)
    W = np.exp(-sign * 2j * np.pi * np.outer(k, k) / r)
    return W @ x


def stockham(x, sign=1):
    """Unnormalised DFT (sign=+1: e^{-i..}, -1: e^{+i..}) of a 1-D complex array."""
    N = x.shape[0]
    radices = factorize(N)
    a = x.astype(np.complex128).copy()
    Ns = 1
    for r in radices:
        b = np.empty_like(a)
        m = N // r
        for j in range(m):
            k = j % Ns
            t = np.arange(r)
            tw = np.exp(-sign * 2j * np.pi * t * k / (Ns * r))
            y = dft_small(a[j + t * m] * tw, sign)
            base = (j - k) * r + k
            b[base + t * Ns] = y
        a = b
        Ns *= r
    return a


def r2c(x):
    """Real length-N (even) -> N/2+1 complex via one length-N/2 complex FFT."""
    N = x.shape[0]
    M = N // 2
    Z = stockham(x[0::2] + 1j * x[1::2], 1)
    k = np.arange(M + 1)
    Zk = Z[k % M]
    Zc = np.conj(Z[(M - k) % M])
    A, B = 0.5 * (Zk + Zc), 0.5 * (Zk - Zc)
    phi = 2 * np.pi * k / N
    return A + (-np.sin(phi) - 1j * np.cos(phi)) * B


def c2r(X, N):
    """N/2+1 complex (Hermitian half) -> real length N, unnormalised (== N * irfft)."""
    M = N // 2
    k = np.arange(M)
    Xk = X[k]
    Xc = np.conj(X[M - k])
    phi = 2 * np.pi * k / N
    Z = (Xk + Xc) + 1j * (np.cos(phi) + 1j * np.sin(phi)) * (Xk - Xc)
    z = stockham(Z, -1)
    out = np.empty(N)
    out[0::2] = z.real
    out[1::2] = z.imag
    return out
