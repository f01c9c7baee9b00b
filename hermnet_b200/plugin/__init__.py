from .calculator import NNCalculator, build_graph  # noqa: F401
