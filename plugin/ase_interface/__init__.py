"""Same import path as the reference (``/root/reference/plugin/ase_interface/__init__.py``)."""
from .calculator import NNCalculator  # noqa: F401
