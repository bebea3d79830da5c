from .activation import *
from .conv import *
from .count import *
from .devoxelize import *
from .downsample import *
from .hash import *
from .query import *
from .voxelize import *
