from .bit_twiddling import *  # noqa: F401,F403
from .comparison import *  # noqa: F401,F403
from .floating import *  # noqa: F401,F403
from .math import *  # noqa: F401,F403
from .trigonometric import *  # noqa: F401,F403
from .ufunc import ufunc  # noqa: F401
