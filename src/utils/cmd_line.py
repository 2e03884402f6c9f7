from lipreading_b200.cli import parseArgsForClassOrScript, build_parser, read_config  # noqa: F401
