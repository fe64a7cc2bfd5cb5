"""link_b200: B200-native (sm_100a) implementation of the LinK hot path behind the reference's
own SparseTensor / nn.Module surface.  `import link_b200 as torchsparse` is the intended drop-in
(see link_b200.compat for the sys.modules shim and INTEGRATION.md)."""
from .operators import *  # noqa: F401,F403
from .tensor import *  # noqa: F401,F403

__version__ = '0.1.0'
