"""Locates and loads a module of the reference checkout (by file path, under a private name) so a drop-in module that
SHADOWS it can re-export everything it does not replace.  The checkout is ``$MOFO_REFERENCE_DIR`` or the first
``sys.path`` entry (other than this directory) that holds the file."""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def find(filename):
    cands = [os.environ.get("MOFO_REFERENCE_DIR")] + list(sys.path)
    for d in cands:
        if not d:
            continue
        d = os.path.abspath(d)
        if d == HERE:
            continue
        p = os.path.join(d, filename)
        if os.path.isfile(p):
            return p
    return None


def load(name):
    """Returns the reference module ``name`` (e.g. "utils") loaded as ``_mofo_reference_<name>``, or None."""
    alias = "_mofo_reference_" + name
    if alias in sys.modules:
        return sys.modules[alias]
    path = find(name + ".py")
    if path is None:
        return None
    spec = importlib.util.spec_from_file_location(alias, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[alias] = mod
    try:
        spec.loader.exec_module(mod)
    except Exception:
        del sys.modules[alias]
        raise
    return mod


def reexport(mod, into):
    if mod is None:
        return
    for k, v in vars(mod).items():
        if not k.startswith("__"):
            into.setdefault(k, v)
