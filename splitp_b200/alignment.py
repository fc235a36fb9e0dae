"""`Alignment` mapping of the drop-in surface (reference: splitp/alignment.py:4-31).

A `UserDict` of {pattern: value} with a `.taxa` tuple.  `sub_alignment` marginalises onto a subset
of taxa (values of patterns that collapse onto the same sub-pattern are ADDED, alignment.py:26-28)
and memoises the result per taxa tuple sorted in parent order (alignment.py:11,30).  The
marginalisation runs on the device through the same flattening scatter as everything else: the
sub-pattern is the base-4 "row index" over the kept taxa.
"""
from collections import UserDict


class Alignment(UserDict):
    def __init__(self, data, taxa):
        self.data = data
        self.taxa = tuple(taxa)
        self.sub_alignments = dict()

    def sub_alignment(self, sub_taxa):
        sub_taxa = tuple(sorted(sub_taxa, key=self.taxa.index))  # ValueError for a foreign taxon, as the reference
        try:
            return self.sub_alignments[sub_taxa]
        except KeyError:
            pass
        if not (set(sub_taxa) <= set(self.taxa)):
            raise ValueError("Sub taxa must be a  subset of the taxa")
        from . import engine
        table = engine.table_from_mapping(self)
        idx = [self.taxa.index(t) for t in sub_taxa]
        sub = engine.marginalise(table, idx)
        result = Alignment(engine.table_to_dict(sub), sub_taxa)
        self.sub_alignments[sub_taxa] = result
        return result
