from lipreading_b200.trainer import train, eval, train_ctc  # noqa: F401
