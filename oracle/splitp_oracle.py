"""CPU oracle for the SplitP hot path -- TEST INFRASTRUCTURE ONLY.

This module is a numpy restatement of the reference algorithm (js51/SplitP v0.3.2) for the
path `pattern counting -> flattening / subflattening -> split score`.  It is the *checker* for the
CUDA path.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import it.  Nothing under `splitp_b200/` imports it, and the product path
has no CPU fallback.

Parity status: PINNED.  `tests/test_oracle_golden.py` checks every function here against
`tests/golden/*.npz|json`, which were produced by importing the unmodified reference from
`/root/reference` (script: `tests/golden/make_golden.py`) and include the reference's own golden
vectors (`tests/test_constructions.py:5-107`, `tests/test_parsers.py:16-26`).
The one exception is the singular-value arithmetic: it lives in scipy (LAPACK dgesdd, unpinned by
`requirements.txt:1-3`), which this oracle calls the same way the reference does
(`splitp/phylogenetics.py:280-300`); no reference test pins `split_score`, so score parity is pinned
only by reference outputs generated in the build container (golden `scores`).

All citations are `file:line` relative to the reference checkout.
"""
from __future__ import annotations

import itertools
from math import floor

import numpy as np
import scipy.linalg

# splitp/constants.py:7-8 -- state order fixes the 2-bit code A=0, C=1, G=2, T=3
STATES = "ACGT"
_CODE = np.full(256, 255, dtype=np.uint8)
for _i, _c in enumerate(STATES):
    _CODE[ord(_c)] = _i
    _CODE[ord(_c.lower())] = _i  # parsers/fasta.py:54 upper-cases each site pattern

# Sign table of the subflattening (splitp/constructions.py:143-161): sign(label, data) = +1 if the
# pair is in `banned`, else -1.  Rows = label state (A,C,G,T), cols = data state (A,C,G,T).
H4 = np.array(
    [[1, -1, -1, 1],
     [1, 1, -1, -1],
     [1, -1, 1, -1],
     [1, 1, 1, 1]], dtype=np.int64)


# ----------------------------------------------------------------------------------------------
# pattern counting  (splitp/parsers/fasta.py:48-70, splitp/simulation.py:43-54)
# ----------------------------------------------------------------------------------------------
def sequences_to_codes(sequences):
    """list of equal-length str/bytes -> uint8 [n, N] with 0..3 for ACGT (any case), 255 otherwise."""
    rows = []
    for s in sequences:
        if isinstance(s, str):
            s = s.encode("latin-1", errors="replace")
        rows.append(_CODE[np.frombuffer(s, dtype=np.uint8)])
    return np.stack(rows) if rows else np.zeros((0, 0), np.uint8)


def codes_to_keys(codes):
    """uint8 [n, N] in 0..3 -> python-int-safe keys.  Key = base-4 number, taxon 0 most significant
    (same significance as `__index_of`, constructions.py:166-171).  n <= 31 -> uint64 array."""
    n = codes.shape[0]
    if n > 31:
        raise ValueError("oracle keys are uint64: n <= 31")
    keys = np.zeros(codes.shape[1], dtype=np.uint64)
    for j in range(n):
        keys = (keys << np.uint64(2)) | codes[j].astype(np.uint64)
    return keys


def keys_to_patterns(keys, n):
    out = []
    for k in keys:
        k = int(k)
        out.append("".join(STATES[(k >> (2 * (n - 1 - j))) & 3] for j in range(n)))
    return out


def patterns_to_keys(patterns):
    n = len(patterns[0]) if len(patterns) else 0
    arr = np.frombuffer("".join(patterns).encode("ascii"), dtype=np.uint8).reshape(len(patterns), n)
    codes = _CODE[arr]
    if (codes > 3).any():
        raise KeyError("non-ACGT character in pattern")  # constructions.py:170 raises KeyError
    return codes_to_keys(codes.T), n


def get_pattern_counts_arrays(codes):
    """parsers/fasta.py:48-63 on a code matrix.  Returns (keys, counts, usable_len) with keys in
    first-occurrence order (the reference's dict insertion order)."""
    valid = (codes <= 3).all(axis=0)  # fasta.py:55-57: site kept only if every char in ACGT
    usable = int(valid.sum())
    keys = codes_to_keys(np.where(codes <= 3, codes, 0)[:, valid])
    uniq, first, counts = np.unique(keys, return_index=True, return_counts=True)
    order = np.argsort(first, kind="stable")
    return uniq[order], counts[order].astype(np.int64), usable


def get_pattern_counts(sequences):
    """Drop-in shape of parsers/fasta.py:48-63: ({pattern: count}, usable_len)."""
    codes = sequences_to_codes(sequences)
    keys, counts, usable = get_pattern_counts_arrays(codes)
    pats = keys_to_patterns(keys, codes.shape[0])
    return {p: int(c) for p, c in zip(pats, counts)}, usable


def get_pattern_counts_wide(codes):
    """parsers/fasta.py:48-63 for any number of taxa (no uint64 key): {pattern: count} in first-occurrence order,
    usable length.  Patterns are compared as byte rows."""
    valid = (codes <= 3).all(axis=0)
    rows = np.ascontiguousarray(codes[:, valid].T)
    uniq, first, counts = np.unique(rows, axis=0, return_index=True, return_counts=True)
    order = np.argsort(first, kind="stable")
    lut = np.frombuffer(STATES.encode(), dtype=np.uint8)
    pats = [lut[r].tobytes().decode("ascii") for r in uniq[order]]
    return dict(zip(pats, counts[order].tolist())), int(valid.sum())


def flattening_reduced_from_dict(pattern_values, idx_a, idx_b):
    """constructions.py:31-55 on a {pattern: value} dict for any number of taxa (row / column indices as python ints:
    the reference's own arithmetic, __index_of :166-171)."""
    def index_of(pat, idx):
        v = 0
        for t in idx:
            v = v * 4 + STATES.index(pat[t])
        return v
    cells = {}
    for pat, val in pattern_values.items():
        cells[(index_of(pat, idx_a), index_of(pat, idx_b))] = val  # assignment: the last pattern wins a cell (:43)
    rows = sorted({r for r, _ in cells})
    cols = sorted({c for _, c in cells})
    ri, ci = {r: i for i, r in enumerate(rows)}, {c: i for i, c in enumerate(cols)}
    out = np.zeros((len(rows), len(cols)))
    for (r, c), v in cells.items():
        out[ri[r], ci[c]] = v
    return out


def pattern_counts_to_probs(patterns, seq_len):
    """parsers/fasta.py:66-70 -- one IEEE division per pattern."""
    return {k: v / seq_len for k, v in patterns.items()}


def sorted_probs_from_counts(counts_dict, sequence_length):
    """simulation.py:50-54 -- lexicographic A<C<G<T key order, value / float(sequence_length)."""
    out = {}
    for k in sorted(counts_dict.keys(), key=lambda p: [STATES.index(c) for c in p]):
        out[k] = counts_dict[k] / float(sequence_length)
    return out


# ----------------------------------------------------------------------------------------------
# splits  (splitp/splits.py:27-59)
# ----------------------------------------------------------------------------------------------
def all_splits(taxa, trivial=False, size=None, string_format=False):
    taxa = list(taxa)
    if string_format and len(taxa) > 35:
        raise ValueError("Cannot generate splits for more than 35 taxa in string format.")
    sizes = [size] if size is not None else list(range(1 if trivial else 2, floor(len(taxa) / 2) + 1))
    for bal in sizes:
        even = bal == len(taxa) / 2
        combos = itertools.combinations(taxa[1:], bal - 1) if even else itertools.combinations(taxa, bal)
        for left in combos:
            if even:
                left = (taxa[0],) + left
            right = tuple(sorted(set(taxa) - set(left), key=taxa.index))
            left = tuple(sorted(left, key=taxa.index))
            if taxa[0] in right:
                left, right = right, left
            if string_format:
                yield f'{"".join(left)}|{"".join(right)}'
            else:
                yield (left, right)


def resolve_split(split, mapping):
    """constructions.py:19-24: '01|23' strings split on '|' and iterated per character; taxa order
    is `mapping.taxa` if present else sorted(union of both sides).  Returns (idxA, idxB) positions
    into the pattern string, in the order given (order = digit significance)."""
    if isinstance(split, str):
        split = split.split("|")
    try:
        taxa = mapping.taxa
    except AttributeError:
        taxa = sorted(set.union(*map(set, split)))
    indexer = {t: i for i, t in enumerate(taxa)}
    return [indexer[s] for s in split[0]], [indexer[s] for s in split[1]]


# ----------------------------------------------------------------------------------------------
# flattenings  (splitp/constructions.py:31-102, 166-171)
# ----------------------------------------------------------------------------------------------
def side_index(keys, n, idx):
    """Base-4 index over the taxa positions `idx`, first listed taxon most significant
    (constructions.py:166-171).  Python ints when the side exceeds 31 taxa are not needed here."""
    out = np.zeros(len(keys), dtype=np.uint64)
    for t in idx:
        d = (keys >> np.uint64(2 * (n - 1 - t))) & np.uint64(3)
        out = (out << np.uint64(2)) | d
    return out


def _last_wins(rows, cols, vals):
    """Assignment semantics (constructions.py:43,101): a later pattern overwrites an earlier one at
    the same (row, col).  Returns de-duplicated (rows, cols, vals) keeping the LAST occurrence."""
    cell = np.stack([rows, cols], axis=1)
    rev = cell[::-1]
    _, first_in_rev = np.unique(rev, axis=0, return_index=True)
    keep = np.sort(len(rows) - 1 - first_in_rev)
    return rows[keep], cols[keep], vals[keep]


def flattening_coo(keys, vals, n, idx_a, idx_b):
    rows = side_index(keys, n, idx_a)
    cols = side_index(keys, n, idx_b)
    if len(idx_a) + len(idx_b) < n or len(set(idx_a) | set(idx_b)) < n:
        rows, cols, vals = _last_wins(rows, cols, np.asarray(vals))
    return rows, cols, np.asarray(vals)


def flattening_dense(keys, vals, n, idx_a, idx_b, dtype=np.float64):
    """`flattening(split, aln, FlatFormat.sparse).todense()` (reference tests densify this way,
    tests/test_constructions.py:41,65,93)."""
    rows, cols, v = flattening_coo(keys, vals, n, idx_a, idx_b)
    out = np.zeros((4 ** len(idx_a), 4 ** len(idx_b)), dtype=dtype)
    out[rows.astype(np.int64), cols.astype(np.int64)] = v
    return out


def flattening_reduced(keys, vals, n, idx_a, idx_b, dtype=np.float64):
    """constructions.py:31-55: rows/cols = sorted unique used indices (zero-valued entries still
    register their row/col, :37-45)."""
    rows, cols, v = flattening_coo(keys, vals, n, idx_a, idx_b)
    ur, ri = np.unique(rows, return_inverse=True)
    uc, ci = np.unique(cols, return_inverse=True)
    out = np.zeros((len(ur), len(uc)), dtype=dtype)
    out[ri, ci] = v
    return out


# ----------------------------------------------------------------------------------------------
# subflattening  (splitp/constructions.py:108-198)
# ----------------------------------------------------------------------------------------------
def pair_tables(keys, vals, n):
    """Joint 4x4 tables N[i, j, x, y] = sum of values of patterns with taxon i in state x and taxon j
    in state y (i == j gives the diagonal marginal table)."""
    vals = np.asarray(vals)
    digits = np.stack([((keys >> np.uint64(2 * (n - 1 - t))) & np.uint64(3)).astype(np.int64) for t in range(n)])
    out = np.zeros((n, n, 4, 4), dtype=vals.dtype if vals.dtype.kind in "iu" else np.float64)
    for i in range(n):
        for j in range(n):
            np.add.at(out[i, j], (digits[i], digits[j]), vals)
    return out


def pair_tables_from_codes(codes, as_counts=False):
    """Same tables as `pair_tables` straight from a code matrix uint8 [n, N] (no pattern keys: works above 31
    taxa).  Only usable sites count (fasta.py:55-57); values are count / usable (fasta.py:66-70) unless as_counts."""
    valid = (codes <= 3).all(axis=0)
    c = codes[:, valid].astype(np.int64)
    n, usable = c.shape
    out = np.zeros((n, n, 4, 4), dtype=np.int64)
    for i in range(n):
        for j in range(n):
            out[i, j] = np.bincount(c[i] * 4 + c[j], minlength=16).reshape(4, 4)
    return (out, usable) if as_counts else (out / float(usable), usable)


def subflattening_from_tables(tables, total, idx_a, idx_b):
    """Entry (3i+c, 3j+d) = (H N_{A_i B_j} H^T)[c, d]; last row/col use the marginals; corner = total
    (derivation: SURVEY.md section 8 row a11; labels from constructions.py:174-189)."""
    a, b = len(idx_a), len(idx_b)
    dt = np.float64 if tables.dtype.kind == "f" else np.int64
    S = np.zeros((3 * a + 1, 3 * b + 1), dtype=dt)
    Hm = H4.astype(dt)
    for i, ta in enumerate(idx_a):
        for j, tb in enumerate(idx_b):
            T = Hm @ tables[ta, tb].astype(dt) @ Hm.T
            S[3 * i:3 * i + 3, 3 * j:3 * j + 3] = T[:3, :3]
    for i, ta in enumerate(idx_a):
        marg = np.diag(tables[ta, ta]).astype(dt)
        S[3 * i:3 * i + 3, 3 * b] = (Hm @ marg)[:3]
    for j, tb in enumerate(idx_b):
        marg = np.diag(tables[tb, tb]).astype(dt)
        S[3 * a, 3 * j:3 * j + 3] = (Hm @ marg)[:3]
    S[3 * a, 3 * b] = total
    return S


def subflattening(keys, vals, n, idx_a, idx_b):
    vals = np.asarray(vals)
    tables = pair_tables(keys, vals, n)
    return subflattening_from_tables(tables, vals.sum(), idx_a, idx_b)


def subflattening_direct(keys, vals, n, idx_a, idx_b):
    """Literal restatement of constructions.py:129-163 (signed sum per label pair, accumulated in
    table order).  O(rows*cols*P*n): small cases only.  Requires the split to cover all taxa, as
    `__reconstruct_pattern` (:192-198) does."""
    a, b = len(idx_a), len(idx_b)
    digits = np.stack([((keys >> np.uint64(2 * (n - 1 - t))) & np.uint64(3)).astype(np.int64) for t in range(n)], axis=1)

    def labels(length):  # :174-189
        for i in range(length):
            for c in range(3):
                lab = [3] * length
                lab[i] = c
                yield lab
        yield [3] * length

    rl, cl = list(labels(a)), list(labels(b))
    out = np.zeros((len(rl), len(cl)))
    for r, row in enumerate(rl):
        for c, col in enumerate(cl):
            pat = np.zeros(n, dtype=np.int64)
            for pos, lab in zip(idx_a, row):
                pat[pos] = lab
            for pos, lab in zip(idx_b, col):
                pat[pos] = lab
            signs = H4[pat[None, :], digits].prod(axis=1)
            acc = 0
            for s, v in zip(signs, vals):  # :152-161, accumulation in table order
                acc += int(s) * v
            out[r, c] = acc
    return out


# ----------------------------------------------------------------------------------------------
# split score  (splitp/phylogenetics.py:280-328)
# ----------------------------------------------------------------------------------------------
def split_score(matrix):
    """phylogenetics.py:280-300: LAPACK singular values, K=4 hard-coded, radicand NOT clamped."""
    m = np.array(matrix.todense() if hasattr(matrix, "todense") else matrix, dtype=np.float64)
    sv = list(scipy.linalg.svd(m, full_matrices=False, check_finite=False, compute_uv=False))
    min_shape = min(m.shape)
    with np.errstate(invalid="ignore"):
        return (1 - (sum(v ** 2 for v in sv[0:4]) / sum(v ** 2 for v in sv[0:min_shape]))) ** (1 / 2)


def split_score_gram(matrix):
    """Independent cross-check used by tests: eigenvalues of the smaller Gram matrix, tail summed
    directly (no cancellation).  Not the reference formula; agrees with it to its noise floor."""
    m = np.asarray(matrix, dtype=np.float64)
    if m.shape[0] > m.shape[1]:
        m = m.T
    lam = np.sort(np.clip(np.linalg.eigvalsh(m @ m.T), 0, None))[::-1]
    tot = lam.sum()
    return float(np.sqrt(lam[4:].sum() / tot)) if tot > 0 else float("nan")


# ----------------------------------------------------------------------------------------------
# banned-pattern flattening, rank-1 / rank-k approximations, rank-1 divergence
# (splitp/constructions.py:94-101, splitp/phylogenetics.py:331-373)
# ----------------------------------------------------------------------------------------------
def digit_count(index, digits, code):
    """How many of the `digits` base-4 digits of `index` equal `code` (str.count of constructions.py:95-98)."""
    index = np.asarray(index, dtype=np.uint64)
    out = np.zeros(index.shape, dtype=np.int64)
    for d in range(digits):
        out += ((index >> np.uint64(2 * d)) & np.uint64(3)) == np.uint64(code)
    return out


def flattening_coo_banned(keys, vals, n, idx_a, idx_b, ban_row=None, ban_col=None):
    """constructions.py:86-102 with ban_row_patterns / ban_col_patterns: an entry whose row (col) pattern holds the
    banned character more than once is assigned 0.  ban_* are state codes 0..3 or None.  Zero entries are dropped
    (a DOK matrix stores nothing for an assigned zero)."""
    rows = side_index(keys, n, idx_a)
    cols = side_index(keys, n, idx_b)
    v = np.array(vals, dtype=np.float64)
    if ban_row is not None:
        v[digit_count(rows, len(idx_a), ban_row) > 1] = 0.0
    if ban_col is not None:
        v[digit_count(cols, len(idx_b), ban_col) > 1] = 0.0
    if len(set(idx_a) | set(idx_b)) < n:
        rows, cols, v = _last_wins(rows, cols, v)
    nz = v != 0
    return rows[nz], cols[nz], v[nz]


def rank_1_vectors(matrix):
    """phylogenetics.py:334-335: r = sum(flattening) = column sums, c = sum(flattening.T) = row sums, both by
    sequential addition of the rows (Python's builtin sum)."""
    m = np.asarray(matrix, dtype=np.float64)
    r = np.zeros(m.shape[1])
    for row in m:
        r = r + row
    c = np.zeros(m.shape[0])
    for col in m.T:
        c = c + col
    return r, c


def rank_1_approximation(matrix):
    """phylogenetics.py:336: r.T @ c -- note the orientation, (#cols x #rows)."""
    r, c = rank_1_vectors(matrix)
    return np.outer(r, c)


def rank_1_divergence(matrix):
    """phylogenetics.py:364-373: sum over non-zero cells of F[x,y] * log(F[x,y] / (r[y] * c[x])), row-major order."""
    m = np.asarray(matrix, dtype=np.float64)
    r, c = rank_1_vectors(m)
    total = 0.0
    for x in range(m.shape[0]):
        nz = np.nonzero(m[x])[0]
        for y in nz:
            total += m[x, y] * np.log(m[x, y] / (r[y] * c[x]))
    return total


def rank_k_approximation(keys, vals, n, idx_a, idx_b):
    """phylogenetics.py:344-361: sum over the four states of (column sums with that state banned on the rows)^T
    times (row sums with that state banned on the columns): a (4^b x 4^a) matrix."""
    R, C = 4 ** len(idx_a), 4 ** len(idx_b)
    out = np.zeros((C, R))
    for code in range(4):
        rows, cols, v = flattening_coo_banned(keys, vals, n, idx_a, idx_b, ban_row=code)
        colsum = np.zeros(C)
        np.add.at(colsum, cols.astype(np.int64), v)
        rows, cols, v = flattening_coo_banned(keys, vals, n, idx_a, idx_b, ban_col=code)
        rowsum = np.zeros(R)
        np.add.at(rowsum, rows.astype(np.int64), v)
        out += np.outer(colsum, rowsum)
    return out


# ----------------------------------------------------------------------------------------------
# Alignment.sub_alignment  (splitp/alignment.py:10-31)
# ----------------------------------------------------------------------------------------------
def sub_alignment(keys, vals, n, sub_idx):
    sub_idx = sorted(sub_idx)
    sub = side_index(keys, n, sub_idx)
    uniq, inv = np.unique(sub, return_inverse=True)
    vals = np.asarray(vals)
    out = np.zeros(len(uniq), dtype=vals.dtype)
    np.add.at(out, inv, vals)
    return uniq, out
