"""Drop-in `flattening` / `subflattening` (reference: splitp/constructions.py:7-198).

Same names, arguments, return types and error behaviour as the reference; the work runs in the
sm_100a kernels behind `engine` (flattening scatter = csrc/flatten.cu, subflattening = csrc/pairs.cu).
"""
from __future__ import annotations

import numpy as np
from scipy.sparse import coo_matrix, dok_matrix

from . import constants, engine
from .enums import FlatFormat


def _taxa_of(split, pattern_probabilities):
    # constructions.py:21-24: the mapping's taxa, else the sorted union of both sides
    try:
        return pattern_probabilities.taxa
    except AttributeError:
        return sorted(set.union(*map(set, split)))


def flattening(split, pattern_probabilities, flattening_format=FlatFormat.sparse):
    """Flattening of `split` ("01|23" or a pair of taxon iterables) from a {pattern: value} mapping.

    FlatFormat.sparse  -> scipy.sparse.dok_matrix (4^|A| x 4^|B|), constructions.py:58-102
    FlatFormat.reduced -> ndarray with all-zero rows/cols dropped, sorted, constructions.py:31-55
    FlatFormat.dense   -> ndarray 4^|A| x 4^|B| (= sparse.todense(); new in this engine)
    Any other format returns None, like the reference (constructions.py:25-28).
    A `engine.PatternTable` is accepted in place of the mapping (device-resident table).
    """
    if isinstance(split, str):
        split = split.split("|")
    if flattening_format not in (FlatFormat.sparse, FlatFormat.reduced, FlatFormat.dense):
        return None
    split = (list(split[0]), list(split[1]))
    taxa = _taxa_of(split, pattern_probabilities)
    idx_a, idx_b = engine.split_positions(split, taxa)  # KeyError for an unknown taxon, as the reference
    table = engine.table_from_mapping(pattern_probabilities)
    if table.num and max(idx_a + idx_b, default=-1) >= table.n:
        raise IndexError("string index out of range")  # pattern[taxa_indexer[s]] in the reference
    if flattening_format is FlatFormat.reduced:
        return engine.flatten_reduced(table, idx_a, idx_b).cpu().numpy()
    if flattening_format is FlatFormat.dense:
        return engine.flatten_dense(table, idx_a, idx_b).cpu().numpy()
    return _sparse(table, idx_a, idx_b)


def _ban_code(char):
    if char is None:
        return -1
    if not isinstance(char, str) or len(char) != 1:
        raise NotImplementedError("banned patterns: a single state character is supported (the only use in the reference)")
    return constants.DNA_state_space.index(char) if char in constants.DNA_state_space else -1  # other chars never occur


def _sparse(table, idx_a, idx_b, ban_row=-1, ban_col=-1):
    # rows / cols (and the banned flags) from the device, assembled into the DOK container the reference returns
    a, b = len(idx_a), len(idx_b)
    if a > 31 or b > 31:
        raise NotImplementedError("sparse flattening: sides are limited to 31 taxa (int64 indices)")
    vals = table.values_f64().cpu().numpy()
    if ban_row >= 0 or ban_col >= 0:
        rows_t, cols_t, banned_t = engine.flatten_coo_banned(table, idx_a, idx_b, ban_row, ban_col)
        vals = np.where(banned_t.cpu().numpy() != 0, 0.0, vals)  # constructions.py:94-99: the entry is assigned 0
    else:
        rows_t, cols_t = engine.flatten_coo(table, idx_a, idx_b)
    rows, cols = rows_t.cpu().numpy(), cols_t.cpu().numpy()
    shape = (4 ** a, 4 ** b)
    if not engine.covers_all(table.n, idx_a, idx_b) and len(rows):
        # assignment semantics (constructions.py:101): the last pattern in table order wins a cell
        cell = rows.astype(object) * shape[1] + cols.astype(object) if shape[1] > 2 ** 31 else rows * shape[1] + cols
        _, first_rev = np.unique(cell[::-1], return_index=True)
        keep = np.sort(len(cell) - 1 - first_rev)
        rows, cols, vals = rows[keep], cols[keep], vals[keep]
    out = dok_matrix(shape)
    if len(rows):
        nz = vals != 0  # assigning 0 to a DOK cell stores nothing
        out = coo_matrix((vals[nz], (rows[nz], cols[nz])), shape=shape).todok()
    return out


def sparse_flattening_with_banned_patterns(split, pattern_probabilities, taxa, ban_row_patterns=None, ban_col_patterns=None):
    """Sparse flattening in which an entry whose row (column) pattern contains the character `ban_row_patterns`
    (`ban_col_patterns`) more than once is assigned 0 (reference: constructions.py:58-105; the only caller,
    phylogenetics.py:344-361, passes single state characters).  `taxa` is the taxon order of the patterns."""
    if isinstance(split, str):
        split = split.split("|")
    idx_a, idx_b = engine.split_positions((list(split[0]), list(split[1])), taxa)
    table = engine.table_from_mapping(pattern_probabilities)
    if table.num and max(idx_a + idx_b, default=-1) >= table.n:
        raise IndexError("string index out of range")
    return _sparse(table, idx_a, idx_b, _ban_code(ban_row_patterns), _ban_code(ban_col_patterns))


def _labels(length):
    # constructions.py:174-189: for each position, A/C/G at that position and T elsewhere; then all-T
    for i in range(length):
        for c in constants.DNA_state_space[:-1]:
            yield "T" * i + c + "T" * (length - i - 1)
    yield "T" * length


def subflattening(split, pattern_probabilities, data=None):
    """Signed-sum subflattening, (3|A|+1) x (3|B|+1) float64 ndarray (constructions.py:108-163).

    Computed as gathers of H N_ij H^T over the pairwise 4x4 joint tables N_ij --
    the same numbers as the reference's O(rows*cols*P*n) signed sums.  `data` is the reference's
    memo dict: its "labels" entry is filled the same way; "coeffs" stays empty because no per-pattern
    sign products are needed.
    """
    # the reference derives the taxa BEFORE splitting a string split (constructions.py:114-117 vs
    # :129-130), so "01|23" with a plain dict counts '|' as a taxon and later fails: reproduced below.
    taxa = _taxa_of(split, pattern_probabilities)
    if data is None:
        data = {}
    data.setdefault("coeffs", {})
    labels = data.setdefault("labels", {})
    if isinstance(split, str):
        split = split.split("|")
    split = (list(split[0]), list(split[1]))
    for side in split:
        if len(side) not in labels:
            labels[len(side)] = list(_labels(len(side)))
    idx_a, idx_b = engine.split_positions(split, taxa)
    n_eff = len(taxa)
    covered = set(idx_a) | set(idx_b)
    for i in range(n_eff):
        if i not in covered:
            raise KeyError(i)  # __reconstruct_pattern (constructions.py:192-198) indexes every position
    table = engine.table_from_mapping(pattern_probabilities)
    if table.num and table.n != n_eff:
        raise IndexError("pattern length does not match the number of taxa")
    pt = engine.pair_tables_from_table(table)
    return engine.subflatten(pt, idx_a, idx_b).cpu().numpy()
