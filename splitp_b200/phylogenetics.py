"""Drop-in `split_score` (reference: splitp/phylogenetics.py:280-328).

    score = sqrt(1 - sum_{i<4} sigma_i^2 / sum_i sigma_i^2),  sigma = singular values of the matrix

sigma_i^2 are the eigenvalues of the Gram matrix on the short side, computed on the device
(fp64 Gram kernel + shared-memory Jacobi / block-Krylov Rayleigh-Ritz, csrc/score.cu).  K = 4 is
hard-coded as in the reference.  Tolerance against the reference's LAPACK path: 1e-9 relative, or the
conditioning floor c*eps/score^2 when that is larger ("score conditioning").
"""
from __future__ import annotations

import numpy as np
import scipy.sparse
import torch

from . import engine


def is_sparse(matrix):
    return scipy.sparse.issparse(matrix)


def split_score(matrix, return_singular_values=False, force_frob_norm_on_dense=False, data_table_for_frob_norm=None):
    """Split score of a flattening / subflattening.  The three keyword arguments are accepted for
    signature compatibility; the reference ignores them too (phylogenetics.py:321-328 passes them to
    parameters that are never read or are unreachable)."""
    if isinstance(matrix, torch.Tensor):
        return engine.score_matrix(matrix.to(engine.device(), torch.float64))[0]
    if is_sparse(matrix):
        if min(matrix.shape) <= 4:
            # scipy.sparse.linalg.svds(k=4) in the reference (phylogenetics.py:306) refuses these
            raise ValueError(f"`k` must be an integer satisfying `0 < k < min(A.shape)`: min(shape)={min(matrix.shape)}")
        coo = matrix.tocoo()
        ur, ri = np.unique(coo.row, return_inverse=True)
        uc, ci = np.unique(coo.col, return_inverse=True)
        dev = engine.device()
        dense = torch.zeros((len(ur), len(uc)), dtype=torch.float64, device=dev)
        if len(coo.data):
            dense[torch.from_numpy(ri).to(dev), torch.from_numpy(ci).to(dev)] = torch.from_numpy(
                np.asarray(coo.data, dtype=np.float64)).to(dev)
        if min(dense.shape) <= 4:
            return 0.0
        return float(engine.score_matrix(dense)[0].item())  # the sparse path returns a Python float (:312)
    m = np.array(matrix, dtype=np.float64)
    if m.ndim != 2:
        raise ValueError("expected a matrix")
    if min(m.shape) == 0:
        return np.float64("nan")
    dev = engine.device()
    return np.float64(engine.score_matrix(torch.from_numpy(np.ascontiguousarray(m)).to(dev))[0].item())
