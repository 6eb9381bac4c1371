"""Drop-in alias: ``import deep_gemm`` resolves to the B200-native implementation in ``deepgemm_b200``."""
import sys as _sys

import deepgemm_b200 as _impl
from deepgemm_b200 import *  # noqa: F401,F403
from deepgemm_b200 import testing, utils  # noqa: F401

for _name in dir(_impl):
    if not _name.startswith('__'):
        globals()[_name] = getattr(_impl, _name)
_sys.modules[__name__ + '.utils'] = utils
_sys.modules[__name__ + '.testing'] = testing
_sys.modules[__name__ + '.utils.math'] = _impl.utils.math
_sys.modules[__name__ + '.utils.layout'] = _impl.utils.layout
__version__ = _impl.__version__
