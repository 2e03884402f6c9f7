from lipreading_b200.data import *  # noqa: F401,F403
from lipreading_b200.data import _collate_fn, FrameCaptionDataset  # noqa: F401
from lipreading_b200.vocab import BOS, EOS, PAD, UNK, MARKERS2ID as _markers2Id, FALLBACK_LABELS as _labels  # noqa: F401
