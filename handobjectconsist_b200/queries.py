"""Sample-dictionary keys -- the members of /root/reference/meshreg/datasets/queries.py:4-46 that
``warpbranch.forward`` and ``ObjBranch.forward`` read (same names, so a reference batch dict works unchanged)."""
from enum import Enum, auto


class BaseQueries(Enum):
    CAMINTR = auto()
    OBJFACES = auto()
    OBJVERTS3D = auto()
    HANDVERTS3D = auto()
    IMAGE = auto()
    OBJCANVERTS = auto()
    OBJCANCORNERS = auto()
    OBJCORNERS3D = auto()


class TransQueries(Enum):
    CAMINTR = auto()
    OBJVERTS3D = auto()
    HANDVERTS3D = auto()
    IMAGE = auto()
    JITTERMASK = auto()
