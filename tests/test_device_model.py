"""The device algorithm (local-offset reformulation, Parseval Z, padded FFT length) is the same algorithm
as the reference's: check its numpy model against the compiled reference's golden vectors.  CPU only."""
import numpy as np
import pytest

from device_model import device_gradient_rep
from test_oracle_golden import GRAD_CASES


@pytest.mark.parametrize("name", GRAD_CASES)
def test_device_model_matches_reference_golden(golden_gradients, name):
    g = golden_gradients
    dims, df, nterms, ipi, min_int, Z, kl = g[name + "__meta"]
    Y = g[name + "__Y"]
    ref = -g[name + "__dC_rep"]
    for ft, tol_f, tol_z in ((np.float64, 1e-10, 1e-11), (np.float32, 2e-6, 1e-6)):
        F, Zm = device_gradient_rep(Y, int(nterms), ipi, int(min_int), df, ft)
        assert np.linalg.norm(F - ref) / np.linalg.norm(ref) < tol_f
        assert abs(Zm - Z) / Z < tol_z
