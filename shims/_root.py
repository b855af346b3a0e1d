"""Where experiment/ and dataset/ live: the directory of the script being run (the reference computes ROOT_PATH from
its own file locations, i.e. its checkout), or LCN_ROOT_PATH."""
import os
import sys


def root_path():
    env = os.environ.get("LCN_ROOT_PATH")
    if env:
        return env
    main = sys.modules.get("__main__")
    f = getattr(main, "__file__", None) or (sys.argv[0] if sys.argv and sys.argv[0] else None)
    return os.path.dirname(os.path.realpath(f)) if f else os.getcwd()
