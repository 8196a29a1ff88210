from .patches import Patches
from .grid import Grid

__all__ = ["Patches", "Grid"]
