"""Workspace paths and logging with the reference's conventions (src/utils/utility.py:16-65):
everything lives under $LIP_READING_WS_PATH/data/{raw,datasets,pickles,weights}."""
import logging
import os

DEFAULT_VERBOSITY = 4
_LOG_FORMAT = "[%(asctime)s %(levelname)5s %(filename)s %(funcName)s:%(lineno)s] %(message)s"
logging.basicConfig(format=_LOG_FORMAT, datefmt="%Y-%m-%d %H:%M:%S")


def getLogger(name, level=logging.DEBUG, verbosity=DEFAULT_VERBOSITY):
    logger = logging.getLogger(name)
    logger.setLevel(max(level, logging.CRITICAL - 10 * verbosity))
    return logger


def getWsDir():
    path = os.getenv("LIP_READING_WS_PATH")
    assert path is not None, ("Environment variable 'LIP_READING_WS_PATH' not found: "
                              "please check project installation and ~/.bashrc")
    return path


def getRelDataPath(*rel):
    return os.path.join(getWsDir(), "data", *rel)


def getRelRawPath(*rel):
    return getRelDataPath("raw", *rel)


def getRelDatasetsPath(*rel):
    return getRelDataPath("datasets", *rel)


def getRelPicklesPath(*rel):
    return getRelDataPath("pickles", *rel)


def getRelWeightsPath(*rel, use_existing=True):
    """weights/<rel>/<n>: n = 0 when use_existing, else the first run index that does not exist yet."""
    path = getRelDataPath("weights", *rel)
    n = 0
    while not use_existing and os.path.isdir(os.path.join(path, str(n))):
        n += 1
    return os.path.join(path, str(n))


def mkdirP(path):
    os.makedirs(path, exist_ok=True)
