"""Enums of the drop-in surface (reference: splitp/enums.py:1-27).

`FlatFormat.dense` does not exist in the reference; it is defined here as
`flattening(split, aln, FlatFormat.sparse).todense()` (how the reference's own tests densify,
tests/test_constructions.py:41,65,93) returned as a C-contiguous float64 ndarray.
"""
from enum import Enum, auto


class _NameEnum(Enum):
    def _generate_next_value_(name, start, count, last_values):  # noqa: N805
        return name


class FlatFormat(_NameEnum):
    sparse = auto()
    reduced = auto()
    dense = auto()


class Method(_NameEnum):
    flattening = auto()
    subflattening = auto()
    distance = auto()
    mutual_information = auto()
