"""Drop-in FASTA helpers (reference: splitp/parsers/fasta.py:3-80).

Reading the text file stays on the host (out of scope); the per-site work --
upper-casing, the A/C/G/T validity rule (fasta.py:54-57), packing and counting -- runs on the device
(csrc/pack.cu, csrc/count.cu).
"""
from __future__ import annotations

import collections

import numpy as np
import torch

from .. import engine


def fasta_to_dict(path_to_file):
    alignment = collections.OrderedDict()
    current = ""
    with open(path_to_file, "r") as handle:
        for line in handle:
            if ">" in line:  # any line containing '>' is a header; the name keeps leading spaces
                current = line.replace(">", "").replace("\n", "")
            elif current in alignment:
                alignment[current] += line.replace("\n", "")
            else:
                alignment[current] = line.replace("\n", "")
    if len(set(len(v) for v in alignment.values())) != 1:
        raise ValueError("Invalid alignment")
    return alignment


read_alignment_from_file = fasta_to_dict


def _to_bytes_matrix(sequences):
    rows = [np.frombuffer(s.encode("latin-1", errors="replace") if isinstance(s, str) else bytes(s), dtype=np.uint8)
            for s in sequences]
    if len(set(len(r) for r in rows)) > 1:
        raise IndexError("string index out of range")  # the reference indexes s[i] up to len(sequences[0])
    return np.stack(rows) if rows else np.zeros((0, 0), np.uint8)


def device_pattern_table(alignment, want_first=True):
    """OrderedDict {taxon: sequence} -> engine.PatternTable (counts, first-site index, usable length)."""
    chars = _to_bytes_matrix(list(alignment.values()))
    aln = engine.pack(chars, is_ascii=True, taxa=list(alignment.keys()), want_planes=False)
    return engine.count_patterns(aln, want_first=want_first)


def get_pattern_counts(alignment):
    """({pattern: count}, usable_length); dict order = first occurrence along the alignment (fasta.py:48-63).
    Up to 31 taxa the keys are uint64 (csrc/count.cu); 32..64 taxa use the 128-bit table (csrc/wide.cu)."""
    if len(alignment) > 31:
        wide, valid, n, N = engine.pack_wide(_to_bytes_matrix(list(alignment.values())), is_ascii=True)
        table = engine.count_patterns_wide(wide, valid, n, N, want_first=True)
        return table.to_dict(order="first"), int(table.divisor)
    table = device_pattern_table(alignment, want_first=True)
    order = torch.argsort(table.first.to(torch.int64) & 0xFFFFFFFF)
    keys = table.keys[order].cpu().numpy().view(np.uint64)
    counts = table.counts[order].cpu().numpy().view(np.uint32)
    return {p: int(c) for p, c in zip(engine.decode_keys(keys, table.n), counts)}, int(table.divisor)


def pattern_counts_to_probs(patterns, seqLen):
    patterns = patterns.copy()
    for key in patterns.keys():
        patterns[key] = patterns[key] / seqLen
    return patterns


def pattern_probs_from_alignment(path_to_file, return_sequence_length=False):
    alignment = read_alignment_from_file(path_to_file)
    counts, sequence_length = get_pattern_counts(alignment)
    probs = pattern_counts_to_probs(counts, sequence_length)
    return (probs, sequence_length) if return_sequence_length else probs
