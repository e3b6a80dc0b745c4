"""The packed kernel-plane layout (FITSNE_FLAG_KPACK; k_gen_kernels / k_hadamard): all four kernel planes of the
convolution -- Ksq, Kb (even in both lattice offsets), Kgrad_x (odd in the column offset), Kgrad_y (odd in the row
offset) -- ride in ONE complex transform, Z = FFT((Kb + Kgx + Kgy) + i*Ksq), and are separated by parity from the four
mirror points of each frequency.  numpy statement of exactly the index conventions the kernels use.  CPU only."""
import numpy as np
import pytest


def _lattice(M, G, p, h, df):
    c = np.arange(M)
    d = np.where(c < G, c, np.where(c > M - G, c - M, 0))
    valid = (c < G) | (c > M - G)
    return d, valid


@pytest.mark.parametrize("df", [1.0, 0.5])
def test_four_kernels_in_one_transform_2d(df):
    M, G, p, h = 48, 21, 3, 0.31
    d, valid = _lattice(M, G, p, h, df)
    dc, dr = d[None, :] * np.ones((M, 1)), d[:, None] * np.ones((1, M))          # plane index = row * M + column
    V = valid[None, :] & valid[:, None]
    t = 1 + h * h * (dc ** 2 + dr ** 2) / df
    kb, ksq = np.where(V, t ** -df, 0.0), np.where(V, t ** -(df + 1), 0.0)
    kgx, kgy = (dc / p) * ksq, (dr / p) * ksq
    Z = np.fft.fft2((kb + kgx + kgy) + 1j * ksq)
    k1, k2 = np.indices((M, M))
    m1, m2 = (M - k1) % M, (M - k2) % M
    I = Z.imag
    Ie, I1, I2, Im = I, I[k1, m2], I[m1, k2], I[m1, m2]                             # k, (k1,-k2), (-k1,k2), -k
    S, ax, ay = (Ie + I1 + I2 + Im) / 4, (Ie - I1 + I2 - Im) / 4, (Ie + I1 - I2 - Im) / 4
    R = Z.real
    kb_hat = (R + R[k1, m2] + R[m1, k2] + R[m1, m2]) / 4
    for got, plane in ((S, ksq), (kb_hat, kb), (1j * ax, kgx), (1j * ay, kgy)):
        assert np.abs(got - np.fft.fft2(plane)).max() < 1e-11


def test_three_kernels_in_one_transform_1d():
    M, G, p, h = 96, 40, 3, 0.27
    d, valid = _lattice(M, G, p, h, 1.0)
    t = 1 + h * h * d ** 2
    kb, ksq = np.where(valid, 1 / t, 0.0), np.where(valid, t ** -2.0, 0.0)
    kg = (d / p) * ksq
    Z = np.fft.fft(kb + kg + 1j * ksq)
    mm = (M - np.arange(M)) % M
    S, a = (Z.imag + Z.imag[mm]) / 2, (Z.imag - Z.imag[mm]) / 2
    assert np.abs(S - np.fft.fft(ksq)).max() < 1e-12
    assert np.abs((Z.real + Z.real[mm]) / 2 - np.fft.fft(kb)).max() < 1e-12
    assert np.abs(1j * a - np.fft.fft(kg)).max() < 1e-12
