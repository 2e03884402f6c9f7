from lipreading_b200.model import VideoEncoder, CharDecodingStep  # noqa: F401
