"""PYTHONPATH shim: `import torchsparse` -> taseg_b200 (put <repo>/taseg_b200/dropin on PYTHONPATH)."""
import sys

import taseg_b200

taseg_b200.install_as_torchsparse()
sys.modules[__name__] = taseg_b200
