from lipreading_b200.face import PRN  # noqa: F401
