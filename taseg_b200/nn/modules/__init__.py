from .activation import *
from .conv import *
from .norm import *
