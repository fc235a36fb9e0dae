"""Alignment generation (reference: splitp/simulation.py:9-56, splitp/model.py:14-75).

The reference evolves one site at a time through networkx (about 830 sites/s), which
cannot produce the 10^6..10^8-site benchmark inputs.  `simulate_codes` draws all sites at once with torch
on whatever device it is given: uniform root state (simulation.py:28), then for every edge the child
state is drawn from column `parent state` of expm(t Q) (simulation.py:18-19).  It is statistically
equivalent to the reference simulator, not bit-identical; parity is always checked on the SAME
alignment fed to both sides.  `generate_alignment` keeps the reference's signature and return type
(a plain dict of sorted patterns -> count / float(sequence_length), simulation.py:42-56) with the
counting done on the device.
"""
from __future__ import annotations

import numpy as np
import torch
from scipy.linalg import expm

from . import engine


class GTR:
    """General time-reversible rate matrix, normalised to one expected substitution per unit time
    (model.py:24-61)."""

    def __init__(self, equilibrium=(0.25, 0.25, 0.25, 0.25), rates=(1, 1, 1, 1, 1, 1), name="GTR"):
        pi = np.asarray(equilibrium, dtype=np.float64)
        if len(pi) != 4 or len(rates) != 6:
            raise ValueError("Incorrect number of parameters for GTR model.")
        if any(x < 0 for x in list(pi) + list(rates)):
            raise ValueError("All parameters must be positive.")
        Q = np.zeros((4, 4))
        Q[np.triu_indices(4, 1)] = rates
        Q = (Q + Q.T) * (np.tile(pi, (4, 1)) - np.diag(pi))
        Q -= np.diag(Q.sum(axis=1))
        self.rate_matrix = Q / -(pi @ np.diag(Q))
        self.init_dist = pi
        self.name = name

    @classmethod
    def JukesCantor(cls, rate=1):
        return cls(rates=[rate] * 6, name="Jukes-Cantor model")

    def transition_matrix(self, t):
        return expm(t * self.rate_matrix)


def simulate_codes(tree, model, sequence_length, seed=0, device=None):
    """uint8 [n_taxa, sequence_length] of codes 0..3, rows in `tree.taxa` order.  `tree`: a `trees.Tree`, or the
    reference's `Phylogeny` (anything with a `.networkx_graph` carrying `branch_length` node attributes and `.taxa`)."""
    from .trees import as_tree
    tree = as_tree(tree)
    device = torch.device(device) if device is not None else engine.device()
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    N = int(sequence_length)
    states = {0: torch.randint(0, 4, (N,), generator=gen, device=device, dtype=torch.uint8)}
    pending = {}  # node -> number of children still to be drawn (inner states are freed as soon as possible)
    for node in range(1, len(tree.parent)):
        pending[tree.parent[node]] = pending.get(tree.parent[node], 0) + 1
    cache = {}
    chunk = 1 << 24  # bounds the float64 temporaries to a few hundred MB whatever the alignment length
    for node in range(1, len(tree.parent)):  # parents precede their children in the node numbering
        t = tree.branch_length[node]
        if t not in cache:
            M = np.asarray(model.transition_matrix(t), dtype=np.float64)  # column = parent state
            # random.choices(states, weights=M[:, parent]) (simulation.py:18-19) normalises the column; it matters for
            # non-symmetric GTR matrices, whose columns do not sum to 1
            M = M / M.sum(axis=0, keepdims=True)
            cache[t] = torch.from_numpy(np.cumsum(M, axis=0).T.copy()).to(device)  # [parent, cumulative child]
        cdf = cache[t]
        par = tree.parent[node]
        child = torch.empty(N, dtype=torch.uint8, device=device)
        u = torch.rand(N, generator=gen, device=device, dtype=torch.float64) if N <= chunk else None
        for b in range(0, N, chunk):
            e = min(N, b + chunk)
            uu = u[b:e] if u is not None else torch.rand(e - b, generator=gen, device=device, dtype=torch.float64)
            c = cdf[states[par][b:e].to(torch.int64)]  # [len, 4]
            child[b:e] = ((uu >= c[:, 0]).to(torch.uint8) + (uu >= c[:, 1]).to(torch.uint8) + (uu >= c[:, 2]).to(torch.uint8))
        states[node] = child
        pending[par] -= 1
        if pending[par] == 0 and par not in tree.names:
            del states[par]
    row = {name: node for node, name in tree.names.items()}
    return torch.stack([states[row[t]] for t in tree.taxa])


def generate_alignment(tree, model, sequence_length, seed=0):
    """{pattern: count / float(sequence_length)} in lexicographic A<C<G<T order (simulation.py:42-56).  `tree` may be
    the reference's `Phylogeny`, `model` anything with `.transition_matrix(t)` (the reference's `model.GTR` included)."""
    codes = simulate_codes(tree, model, sequence_length, seed)
    if codes.shape[0] > 31:  # 128-bit pattern keys
        wide, valid, n, N = engine.pack_wide(codes)
        wt = engine.count_patterns_wide(wide, valid, n, N)
        wt.divisor = float(sequence_length)
        return wt.to_dict(as_counts=False)
    aln = engine.pack(codes, is_ascii=False, want_planes=False)
    table = engine.count_patterns(aln)
    table.divisor = float(sequence_length)
    return engine.table_to_dict(table)
