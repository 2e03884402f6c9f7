"""Thin stand-ins for the two decoders that FEED the hot path and are out of scope (SURVEY §2 rows
4-5): frame reads (reference: imageio/ffmpeg, src/utils/data/video.py) and WebVTT caption parsing
(reference: pycaption, src/utils/data/caption.py).  Neither wheel is in this image; these keep the
`generate_dataview` CLI runnable on real files when OpenCV is importable."""
import collections
import re
import unicodedata


class VideoReader:
    def __init__(self, vid_path):
        import cv2
        self._cv2 = cv2
        self.cap = cv2.VideoCapture(vid_path)
        assert self.cap.isOpened(), "cannot open " + vid_path

    def get_frame_idx(self, seconds):
        return int(seconds * 29.97)

    def getNumFrames(self):
        return int(self.cap.get(self._cv2.CAP_PROP_FRAME_COUNT))

    def genFrames(self, lo, hi):
        self.cap.set(self._cv2.CAP_PROP_POS_FRAMES, lo)
        out = []
        for _ in range(lo, hi):
            ok, frame = self.cap.read()
            if not ok:
                break
            out.append(self._cv2.cvtColor(frame, self._cv2.COLOR_BGR2RGB))
        return out


_TS = re.compile(r"(\d+):(\d\d):(\d\d)[.,](\d{3})\s*-->\s*(\d+):(\d\d):(\d\d)[.,](\d{3})")
_patterns = (r"\(.*\)", r"<[^>]*>", r"\[.*\]", r"\{.*\}", r"stephen:", r">>")


def extract_captions(cap_path):
    """WebVTT -> OrderedDict {(start_s, end_s): text} (caption.py:37-66)."""
    out = collections.OrderedDict()
    with open(cap_path, encoding="utf-8") as fh:
        blocks = re.split(r"\n\s*\n", fh.read())
    for blk in blocks:
        lines = [ln for ln in blk.strip().splitlines() if ln.strip()]
        for i, ln in enumerate(lines):
            m = _TS.search(ln)
            if m:
                g = [int(x) for x in m.groups()]
                start = g[0] * 3600 + g[1] * 60 + g[2] + g[3] / 1000.0
                end = g[4] * 3600 + g[5] * 60 + g[6] + g[7] / 1000.0
                out[(start, end)] = " ".join(lines[i + 1:])
                break
    return out


def prune_and_filter_captions(captions, patterns=None, conditions=None, union=True):
    """Lower-case, strip speaker tags / bracketed cues, NFKD->ascii, drop captions of <= 2 words
    (caption.py:20-32,68-110)."""
    patterns = patterns or _patterns
    regex = re.compile("|".join(patterns))
    out = collections.OrderedDict()
    for key, cap in captions.items():
        cap = regex.sub("", cap.lower())
        cap = unicodedata.normalize("NFKD", cap).encode("ascii", "ignore").decode("ascii")
        cap = " ".join(cap.split())
        if len(cap.split()) > 2:
            out[key] = cap
    return out
