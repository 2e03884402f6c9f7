"""Vocabulary constants of the reference (src/data/data_loader.py:29-35)."""
BOS, EOS, PAD, UNK = "<BOS>", "<EOS>", "<PAD>", "<UNK>"
MARKERS2ID = {PAD: 0, BOS: 1, EOS: 2, UNK: 3}
ID2MARKERS = {v: k for k, v in MARKERS2ID.items()}
# hard-coded fallback label set used when labels.json cannot be opened (data_loader.py:35,106-108)
FALLBACK_LABELS = [" ", "!", "\"", "#", "$", "%", "&", "'", "(", ")", "*", "+", ",", "-", ".", "/"] + \
    [str(d) for d in range(10)] + [":", ";", "<", ">", "?", "@", "[", "]"] + \
    [chr(c) for c in range(ord("a"), ord("z") + 1)]


def build_char2idx(labels=None):
    char2idx = dict(MARKERS2ID)
    for ch in (FALLBACK_LABELS if labels is None else labels):
        char2idx[ch] = len(char2idx)
    return char2idx
