"""``plugin.ase_interface.calculator`` -- reference import path, B200-native implementation."""
from hermnet_b200.plugin.calculator import NNCalculator, build_graph  # noqa: F401
