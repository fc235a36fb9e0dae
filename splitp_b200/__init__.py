"""splitp_b200 -- B200-native engine for the SplitP hot path, drop-in for
`splitp.flattening / subflattening / split_score / all_splits / FlatFormat / Alignment`
(reference: splitp/__init__.py:2-18)."""
from . import _lib  # noqa: F401  (fails loudly when the sm_100a library is missing)
from . import alignment, constants, constructions, engine, phylogenetics, simulation, splits, trees  # noqa: F401
from .alignment import Alignment  # noqa: F401
from .constructions import flattening, subflattening  # noqa: F401
from .enums import FlatFormat, Method  # noqa: F401
from .phylogenetics import erickson_SVD, split_score  # noqa: F401
from .simulation import generate_alignment  # noqa: F401
from .splits import all_splits  # noqa: F401

__all__ = ["flattening", "subflattening", "split_score", "all_splits", "FlatFormat", "Method", "Alignment",
           "generate_alignment", "erickson_SVD", "engine", "trees", "simulation", "splits", "constructions", "phylogenetics"]
