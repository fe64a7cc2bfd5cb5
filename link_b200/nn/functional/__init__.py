from .conv import *  # noqa: F401,F403
from .count import *  # noqa: F401,F403
from .devoxelize import *  # noqa: F401,F403
from .downsample import *  # noqa: F401,F403
from .hash import *  # noqa: F401,F403
from .norm import *  # noqa: F401,F403
from .query import *  # noqa: F401,F403
from .voxelize import *  # noqa: F401,F403
