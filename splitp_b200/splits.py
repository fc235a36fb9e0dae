"""Split enumeration (reference: splitp/splits.py:7-59).  Host-side only: it defines the work-unit
order and the split representation the engine accepts (2-tuple of taxon tuples, or "012|345")."""
from itertools import combinations
from math import floor

from numpy.random import shuffle


def split_balance(s, asTuple=False):
    left, right = s.split("|")
    return (len(left), len(right)) if asTuple else f"{len(left)}|{len(right)}"


def format_split(tree, split):
    if isinstance(split, str):
        return split
    if len(split[0]) + len(split[1]) > 35:
        raise ValueError("Cannot produce string format for split with more than 35 taxa.")
    if not all(len(taxon) == 1 for taxon in tree.get_taxa()):
        raise ValueError("Cannot produce string format for split with taxa name of length > 1.")
    return f'{"".join(split[0])}|{"".join(split[1])}'


def all_splits(tree, trivial=False, size=None, randomise=False, string_format=False):
    """Generator over the splits of `tree.taxa`: sizes 2..floor(n/2) ascending (1.. with `trivial`),
    `itertools.combinations` order inside a size, balanced splits de-duplicated by pinning taxa[0] to
    the left, taxa[0] always on the left side, both sides in `tree.taxa` order."""
    taxa = tree.taxa
    n = len(taxa)
    if string_format and n > 35:
        raise ValueError("Cannot generate splits for more than 35 taxa in string format. Use string_format=False.")
    sizes = [size] if size is not None else list(range(1 if trivial else 2, floor(n / 2) + 1))
    position = {t: i for i, t in enumerate(taxa)}
    for left_size in sizes:
        balanced = left_size == n / 2
        chosen = combinations(taxa[1:], left_size - 1) if balanced else combinations(taxa, left_size)
        if randomise:
            chosen = list(chosen)
            shuffle(chosen)
        for pick in chosen:
            if balanced:
                pick = (taxa[0],) + tuple(pick)
            inside = set(pick)
            left = tuple(sorted(pick, key=position.__getitem__))
            right = tuple(t for t in taxa if t not in inside)
            if taxa[0] in right:
                left, right = right, left
            yield format_split(tree, (left, right)) if string_format else (left, right)
