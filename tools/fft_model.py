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


# ---------------------------------------------------------------------------------------------- prime-factor (Good-Thomas) stages
# Model of Pfa<> / PStage<> in admp_b200/csrc/fft_fast.cuh: for pairwise-coprime radices R1 x R2 x R3 = N the length-N DFT is the
# R1 x R2 x R3 multi-dimensional DFT of the array w[i1, i2, i3] = x[(i1 S1 + i2 S2 + i3 S3) mod N], S_d = N / R_d (Good's map),
# whose element (k1, k2, k3) is the frequency k = (k1 T1 + k2 T2 + k3 T3) mod N, T_d = S_d * (S_d^-1 mod R_d) (CRT map).
def pfa_constants(radices):
    N = int(np.prod(radices))
    S = [N // r for r in radices]
    T = [(s * pow(s % r, -1, r)) % N if r > 1 else 0 for s, r in zip(S, radices)]
    return N, S, T


def pfa_fft(x, radices, sign=1):
    """DFT by the kernel's stages: gather through Good's map into the row-major work buffer (i1 fastest), one in-place pass per
    dimension, scatter through the CRT map. Returns the natural-order spectrum."""
    R1, R2, R3 = (list(radices) + [1, 1])[:3]
    N, S, T = pfa_constants((R1, R2, R3))
    assert x.shape[0] == N
    work = np.empty(N, dtype=np.complex128)
    for i3 in range(R3):
        for i2 in range(R2):
            base = (i2 * S[1] + i3 * S[2]) % N
            pts = [(base + t * S[0]) % N for t in range(R1)]            # good1 / good1_off
            work[R1 * (i2 + R2 * i3) + np.arange(R1)] = dft_small(x[pts], sign)   # rm1
    for i3 in range(R3):
        for i1 in range(R1):
            pos = i1 + R1 * (np.arange(R2) + R2 * i3)                   # rm2
            work[pos] = dft_small(work[pos], sign)
    out = np.empty(N, dtype=np.complex128)
    for b in range(R1 * R2):                                            # last dimension (3; 2 for two factors is the loop above)
        i1, i2 = b % R1, b // R1
        pos = b + R1 * R2 * np.arange(R3)                               # rml
        v = dft_small(work[pos], sign) if R3 > 1 else work[pos]
        base = (i1 * T[0] + i2 * T[1]) % N
        for t in range(R3):
            out[(base + (t * T[2]) % N) % N] = v[t]                     # crtl / crtl_off
    return out


def pfa_freq_of_position(radices):
    """frequency of every row-major work-buffer position (Pfa::freq) and its inverse (Pfa::pos_of_freq)."""
    R1, R2, R3 = (list(radices) + [1, 1])[:3]
    N, S, T = pfa_constants((R1, R2, R3))
    p = np.arange(N)
    freq = ((p % R1) * T[0] + ((p // R1) % R2) * T[1] + (p // (R1 * R2)) * T[2]) % N
    k = np.arange(N)
    pos = k % R1 + R1 * (k % R2 + R2 * (k % R3 if R3 > 1 else 0))
    return freq, pos
