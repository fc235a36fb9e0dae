"""Split enumeration, host side only (reference interface: splitp/splits.py:7-59).

`all_splits` fixes the work-unit order of the engine and the split representation it accepts: a pair of taxon
tuples, or the "012|345" string form for single-character taxa.  Splits are enumerated on taxon POSITIONS and
only mapped to names when they are yielded.
"""
import itertools
import math

import numpy.random


def split_balance(s, asTuple=False):
    """'AB|CDE' -> '2|3' (or (2, 3))."""
    sizes = tuple(len(part) for part in s.split("|"))
    return sizes if asTuple else "|".join(str(x) for x in sizes)


def format_split(tree, split):
    """Pair of taxon tuples -> 'AB|CDE'; strings pass through.  Needs single-character taxa and at most 35 of them."""
    if isinstance(split, str):
        return split
    left, right = split
    if len(left) + len(right) > 35:
        raise ValueError("string format is limited to splits of at most 35 taxa")
    if any(len(name) != 1 for name in tree.get_taxa()):
        raise ValueError("string format needs taxon names of length 1")
    return "".join(left) + "|" + "".join(right)


def all_splits(tree, trivial=False, size=None, randomise=False, string_format=False):
    """Generator over the splits of `tree.taxa`.

    Order: left-side sizes 2..floor(n/2) ascending (from 1 with `trivial`; only `size` if given), and inside a size
    the order of `itertools.combinations`.  A balanced split (size == n/2) is produced once, with taxon 0 pinned to
    the left; in every split taxon 0 ends up on the left side and both sides keep the order of `tree.taxa`.
    `randomise` shuffles the combinations of each size (numpy's global RNG, like the reference)."""
    names = list(tree.taxa)
    n = len(names)
    if string_format and n > 35:
        raise ValueError("string_format=True supports at most 35 taxa; use string_format=False")
    wanted = [size] if size is not None else range(1 if trivial else 2, math.floor(n / 2) + 1)
    everyone = range(n)
    for k in wanted:
        if 2 * k == n:  # balanced: choose the k - 1 companions of taxon 0
            picks = ((0,) + rest for rest in itertools.combinations(range(1, n), k - 1))
        else:
            picks = itertools.combinations(everyone, k)
        if randomise:
            picks = list(picks)
            numpy.random.shuffle(picks)
        for chosen in picks:
            inside = set(chosen)
            side_a = tuple(names[i] for i in everyone if i in inside)
            side_b = tuple(names[i] for i in everyone if i not in inside)
            if 0 not in inside:
                side_a, side_b = side_b, side_a
            yield format_split(tree, (side_a, side_b)) if string_format else (side_a, side_b)
