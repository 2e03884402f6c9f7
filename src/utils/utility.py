from lipreading_b200.workspace import *  # noqa: F401,F403
