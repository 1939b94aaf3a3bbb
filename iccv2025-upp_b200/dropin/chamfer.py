"""Drop-in top-level `chamfer` module: resolves `import chamfer`
(reference extensions/chamfer_dist/__init__.py:10) so that file runs unmodified."""
from upp_b200.chamfer import backward, forward  # noqa: F401
