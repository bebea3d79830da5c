from .utils import *
