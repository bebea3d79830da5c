from .apply import *
from .kernel import *
