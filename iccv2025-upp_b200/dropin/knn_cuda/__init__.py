"""Drop-in `knn_cuda` package: resolves `from knn_cuda import KNN`
(reference models/Point_MAE_unify.py:16)."""
from upp_b200.knn import KNN  # noqa: F401

__version__ = "0.2+upp_b200"
