"""Minimal offline stand-in for the PyPI package ``scoring-matrices`` (~=0.3).

pytrimal's Cython extension ``cimport``s ``scoring_matrices.lib.ScoringMatrix`` as
the base class of ``pytrimal.SimilarityMatrix`` (src/pytrimal/_trimal.pyx:83,1867).
The real package cannot be installed here (no network, not in the wheelhouse), and
it contributes no arithmetic to the statistics hot path.  This stand-in provides
just the surface pytrimal touches -- ``_size``, ``_matrix``, ``alphabet``, ``name``,
``from_name("BLOSUM62")``, ``shuffle``, ``len``, iteration -- so that the reference's
own extension can be built and driven with ``platform="cuda"``.  It is test
infrastructure for integration/, never shipped as part of the CUDA library.
"""
from .lib import ScoringMatrix

__all__ = ["ScoringMatrix"]
__version__ = "0.3.0+standin"
