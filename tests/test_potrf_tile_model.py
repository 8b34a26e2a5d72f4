"""CPU check of the diagonal-tile kernel's DESIGN: tools/potrf_tile_model.py replays potrf_tile_kernel2's shared-memory index
arithmetic, fragment maps, warp work split and phase order in NumPy (lane by lane) and must reproduce LAPACK's factor and its
inverse; the input's upper triangle is NaN-poisoned, so a read of anything that has not been written shows up in the outputs.
The CUDA kernel itself is tested on the GPU (tests/test_gpu_parity.py); this guards the algorithm where no GPU is available."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import potrf_tile_model as model  # noqa: E402


def test_tile_model_reproduces_cholesky_and_inverse():
    for seed in (0, 3):
        A = model.demo_matrix(seed)
        L, W = model.TileModel(A).run()
        Lr = np.linalg.cholesky(A)
        assert not np.isnan(L).any() and not np.isnan(W).any()
        np.testing.assert_allclose(L, Lr, rtol=0, atol=1e-13)
        np.testing.assert_allclose(W @ Lr, np.eye(model.T), rtol=0, atol=1e-12)
        assert np.all(np.triu(L, 1) == 0.0) and np.all(np.triu(W, 1) == 0.0)
