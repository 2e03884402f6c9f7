from lipreading_b200.face import *  # noqa: F401,F403
from lipreading_b200.face import _applyPadding, _getSharedPrn  # noqa: F401
