"""Index-math model of the custom FFT (tools/fft_model.py) against numpy.fft. CPU only."""
import numpy as np
import pytest

from tools import fft_model as fm


@pytest.mark.parametrize('n', [154, 77, 308, 616, 60, 26, 96])
def test_stockham_matches_numpy(n):
    rng = np.random.default_rng(n)
    x = rng.normal(size=n) + 1j * rng.normal(size=n)
    np.testing.assert_allclose(fm.stockham(x, 1), np.fft.fft(x), atol=1e-11)
    np.testing.assert_allclose(fm.stockham(x, -1), np.fft.ifft(x) * n, atol=1e-11)


@pytest.mark.parametrize('n', [154, 308, 30, 64])
def test_real_packing(n):
    rng = np.random.default_rng(n)
    x = rng.normal(size=n)
    X = fm.r2c(x)
    np.testing.assert_allclose(X, np.fft.rfft(x), atol=1e-11)
    np.testing.assert_allclose(fm.c2r(X, n), x * n, atol=1e-10)


def test_factorize_rejects_unsupported():
    assert fm.factorize(154) == [11, 7, 2]
    assert fm.factorize(1232) == [11, 7, 8, 2] or fm.factorize(1232) == [11, 7, 4, 4]
    assert fm.factorize(17 * 4) is None


@pytest.mark.parametrize('radices', [(11, 7), (11, 14), (11, 7, 4), (11, 7, 8), (11, 7, 16), (3, 5, 4)])
def test_prime_factor_stages_match_numpy(radices):
    """Good's input map + in-place dimensions + CRT output map (admp_b200/csrc/fft_fast.cuh Pfa<> / PStage<>) = the plain DFT, with no
    twiddle factors, for every size of the mesh family."""
    n = int(np.prod(radices))
    rng = np.random.default_rng(n)
    x = rng.normal(size=n) + 1j * rng.normal(size=n)
    np.testing.assert_allclose(fm.pfa_fft(x, radices, 1), np.fft.fft(x), atol=1e-10)
    np.testing.assert_allclose(fm.pfa_fft(x, radices, -1), np.fft.ifft(x) * n, atol=1e-10)
    freq, pos = fm.pfa_freq_of_position(radices)
    assert sorted(freq.tolist()) == list(range(n)) and np.array_equal(freq[pos], np.arange(n)) and freq[0] == 0


def test_prime_factor_round_trip_needs_no_reordering():
    """The fused X pass scales the spectrum in work-buffer order (tables permuted by Pfa::freq) and transforms back through the same
    maps: forward, multiply by g(k), inverse == ifft(g * fft(x))."""
    radices = (11, 7, 4)
    n = 308
    rng = np.random.default_rng(3)
    x = rng.normal(size=n) + 1j * rng.normal(size=n)
    g = rng.normal(size=n)
    X = fm.pfa_fft(x, radices, 1)                  # natural order (the model scatters; the kernel keeps work order + permuted g)
    y = fm.pfa_fft(g * X, radices, -1)
    np.testing.assert_allclose(y, np.fft.ifft(g * np.fft.fft(x)) * n, atol=1e-9)
