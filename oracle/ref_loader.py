"""Import helper for the compiled reference modules in oracle/_ref/.

TEST INFRASTRUCTURE ONLY (tests/, golden-vector generation, bench.py cpu_baseline).

Usage:
    from oracle.ref_loader import load_ref
    ref = load_ref()          # None when oracle/_ref has not been built
    ref.utils.convert_board_to_2d_matrix_POEB(board, 'White')
    ref.board_utils.enumerate_legal_moves(board, 'White')
"""
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")


class _Ref(object):
    pass


_cached = None


def available():
    return os.path.exists(os.path.join(_REF, ".built"))


def load_ref():
    """Returns a namespace with .utils and .board_utils (and .tree_builder etc. on
    request via .module(name)), or None when oracle/_ref is absent."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        return None
    # board_utils.pyx:6-7 imports pandas then bottleneck.argpartition; bottleneck is
    # not installed.  Import pandas FIRST (its optional-dependency probe would trip
    # over a fake bottleneck), then register the stand-in.
    import numpy
    import pandas  # noqa: F401
    if "bottleneck" not in sys.modules:
        bn = types.ModuleType("bottleneck")
        bn.argpartition = numpy.argpartition
        bn.__version__ = "0.0-standin"
        sys.modules["bottleneck"] = bn
    if _REF not in sys.path:
        sys.path.insert(0, _REF)
    ref = _Ref()
    ref.utils = importlib.import_module("tools.utils")
    ref.board_utils = importlib.import_module("Breakthrough_Player.board_utils")
    ref.module = importlib.import_module
    _cached = ref
    return ref
