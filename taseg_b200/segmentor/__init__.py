from .unet import *
from .utils import *
