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
