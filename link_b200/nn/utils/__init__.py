from .apply import *  # noqa: F401,F403
from .kernel import *  # noqa: F401,F403
