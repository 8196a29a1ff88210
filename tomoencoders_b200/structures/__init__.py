from .patches import Patches
from .grid import Grid
from .datafile import DataFile

__all__ = ["Patches", "Grid", "DataFile"]
