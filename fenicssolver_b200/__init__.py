"""B200-native FEM assemble-and-solve behind the FenicsSolver settings API."""
__version__ = "0.1"
