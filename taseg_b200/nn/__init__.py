from .modules import *
from . import functional, utils
