from lipreading_b200.ctc import ctc_loss  # noqa: F401
