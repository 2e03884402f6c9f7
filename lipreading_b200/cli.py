"""Command line derived from a function signature, the reference's flag conventions
(src/utils/cmd_line.py:91-141): bool default -> store_true flag, tuple/list default -> nargs='+',
otherwise type(default); plus -v/--verbosity.  Config files are one `--flag[=value]` per line and are
splatted with `$(cat config/...)`; later flags override earlier ones (README.md:138-165).
(`inspect.getargspec`, which the reference uses, no longer exists in Python 3.11+.)"""
import argparse
import inspect

from .workspace import DEFAULT_VERBOSITY


def build_parser(fn):
    parser = argparse.ArgumentParser(description=(inspect.getdoc(fn) or "").strip() or None)
    for name, prm in inspect.signature(fn).parameters.items():
        if name in ("self", "logger"):
            continue
        if prm.default is inspect.Parameter.empty:
            parser.add_argument("--" + name, default=None, type=str)
        elif isinstance(prm.default, bool):
            parser.add_argument("--" + name, default=prm.default, action="store_true")
        elif isinstance(prm.default, (tuple, list)):
            parser.add_argument("--" + name, default=prm.default, nargs="+", help="Tuple of " + name)
        else:
            parser.add_argument("--" + name, default=prm.default,
                                type=type(prm.default) if prm.default is not None else str)
    parser.add_argument("-v", "--verbosity", default=DEFAULT_VERBOSITY, type=int,
                        help="0 CRITICAL .. 4 DEBUG (default 4)")
    return parser


def parseArgsForClassOrScript(fn, argv=None):
    assert inspect.isfunction(fn) or inspect.ismethod(fn)
    args = build_parser(fn).parse_args(argv)
    if args.verbosity > 0:
        doc = inspect.getdoc(fn)
        assert doc is not None, "Please write documentation :)"
        print("\n" + doc.strip() + "\n\nArguments and corresponding default or set values")
        for k, v in vars(args).items():
            if k != "verbosity":
                print("\t{}={}".format(k, v if v is not None else ""))
        print()
    return args


def read_config(path):
    """One flag per line -> argv list (blank lines ignored)."""
    with open(path) as fh:
        return [ln.strip() for ln in fh if ln.strip()]
