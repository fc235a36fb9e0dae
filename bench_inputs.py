"""Synthetic benchmark inputs shared by both arms of bench.py (numpy / scipy only: no torch, no splitp_b200, no oracle).

The reference's own generator (splitp/simulation.py:9-56) evolves one site at a time through networkx (about 830
sites/s), so neither arm can use it at 10^6 .. 10^8 sites.  `simulate_codes` draws all sites at once: uniform root state
(simulation.py:28), then for every edge the child state from column `parent state` of expm(t Q) (simulation.py:18-19,
model.py:14-16).  Both arms call it with the same seed, so they score the IDENTICAL alignment.

The balanced topology and the leaf names are those of splitp.trees.balanced_newick_tree (trees.py:6-29): two halves per
node, the larger half first on a left child; leaves named 0-9A-Z left to right for n <= 36, t0.. otherwise; taxa sorted
(phylogeny.py:49).
"""
from __future__ import annotations

import itertools

import numpy as np
from scipy.linalg import expm

STATES = "ACGT"


def balanced_parents(num_taxa):
    """(parent list, {node: leaf name}) of the balanced tree; node 0 is the root, parents precede children."""
    if num_taxa % 2 != 0 or num_taxa < 2:
        raise ValueError(f"There is no balanced tree on {num_taxa} taxa. Please specify an even number.")
    parent, names, counter = [-1], {}, [0]

    def leaf_name(i):
        return str(np.base_repr(i, base=max(i + 1, 2))) if num_taxa <= 36 else f"t{i}"

    def grow(node, size, left):
        if size == 1:
            names[node] = leaf_name(counter[0])
            counter[0] += 1
            return
        big = size // 2 + (size % 2 if left else 0)
        for part, is_left in ((big, True), (size - big, False)):
            parent.append(node)
            grow(len(parent) - 1, part, is_left)

    for is_left in (True, False):
        parent.append(0)
        grow(len(parent) - 1, max(num_taxa // 2, 1), is_left)
    return parent, names


def taxa_of(names):
    return sorted(names.values())


def rate_matrix(model):
    """model = "JC" (GTR.JukesCantor(1/2), model.py:66-75) or "GTR" (pi = .1 .2 .3 .4, rates 1..6, model.py:24-61);
    normalised to one expected substitution per unit time."""
    if model == "JC":
        pi, rates = np.full(4, 0.25), [0.5] * 6
    elif model == "GTR":
        pi, rates = np.array([0.1, 0.2, 0.3, 0.4]), [1, 2, 3, 4, 5, 6]
    else:
        raise ValueError(model)
    Q = np.zeros((4, 4))
    Q[np.triu_indices(4, 1)] = rates
    Q = (Q + Q.T) * (np.tile(pi, (4, 1)) - np.diag(pi))
    Q -= np.diag(Q.sum(axis=1))
    return Q / -(pi @ np.diag(Q))


def simulate_codes(num_taxa, branch_length, model, num_sites, seed):
    """uint8 [num_taxa, num_sites] of codes 0..3 (A, C, G, T), rows in sorted-taxa order."""
    parent, names = balanced_parents(num_taxa)
    rng = np.random.default_rng(int(seed))
    N = int(num_sites)
    M = expm(float(branch_length) * rate_matrix(model))  # column = parent state (simulation.py:18-19)
    M = M / M.sum(axis=0, keepdims=True)                # random.choices normalises the column (matters for GTR)
    cdf = np.cumsum(M, axis=0).T.copy()                 # [parent state, cumulative child state]
    pending = {}
    for node in range(1, len(parent)):
        pending[parent[node]] = pending.get(parent[node], 0) + 1
    states = {0: rng.integers(0, 4, size=N, dtype=np.uint8)}
    chunk = 1 << 22
    for node in range(1, len(parent)):
        par = states[parent[node]]
        child = np.empty(N, dtype=np.uint8)
        for b in range(0, N, chunk):
            e = min(N, b + chunk)
            u = rng.random(e - b)
            c = cdf[par[b:e]]
            child[b:e] = (u >= c[:, 0]).astype(np.uint8) + (u >= c[:, 1]) + (u >= c[:, 2])
        states[node] = child
        pending[parent[node]] -= 1
        if pending[parent[node]] == 0 and parent[node] not in names:
            del states[parent[node]]
    row = {name: node for node, name in names.items()}
    return np.stack([states[row[t]] for t in taxa_of(names)])


def codes_to_sequences(codes, taxa):
    """{taxon: ACGT string}: the input type of splitp.parsers.fasta.get_pattern_counts (fasta.py:48)."""
    lut = np.frombuffer(STATES.encode(), dtype=np.uint8)
    return {t: lut[codes[i]].tobytes().decode("ascii") for i, t in enumerate(taxa)}


def all_splits_idx(n, size=None):
    """Position form of splitp.all_splits (splits.py:27-59): sizes 2 .. n // 2 ascending, itertools.combinations order,
    position 0 always on the left, even splits listed once."""
    out = []
    sizes = [size] if size is not None else range(2, n // 2 + 1)
    everyone = list(range(n))
    for a in sizes:
        for left in itertools.combinations(everyone, a):
            right = [t for t in everyone if t not in left]
            if 2 * a == n and 0 not in left:
                continue
            if 0 in left:
                out.append((list(left), right))
            else:
                out.append((right, list(left)))
    return out


def random_splits_idx(n, count, seed=5):
    """`count` random non-trivial splits, side sizes uniform in 2 .. n // 2 (BASELINE configs[4])."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(count):
        a = int(rng.integers(2, n // 2 + 1))
        left = set(rng.choice(n, size=a, replace=False).tolist())
        out.append(([i for i in range(n) if i in left], [i for i in range(n) if i not in left]))
    return out
