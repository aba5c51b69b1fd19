"""spgnn_b200 — B200-native GNN stage of SPGNN (airway-tree labelling): CUDA kernels behind the reference's
Python layer API.  ``import spgnn_b200`` needs the in-tree ``libspgnn_b200.so`` (``python -m spgnn_b200.build``)."""
from ._lib import SpgnnError, lib  # noqa: F401

__version__ = "0.1.0"
