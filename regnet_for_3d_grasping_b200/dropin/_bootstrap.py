"""Puts the repository root on sys.path so the drop-in modules can import the implementation package."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def installed_elsewhere(name):
    """The real third-party package `name`, imported from any sys.path entry other than this drop-in directory, or None.
    The stand-ins next to this file (open3d, transforms3d, tensorboardX) call it first and step aside when the real
    package is installed."""
    import importlib.machinery
    import importlib.util
    here = os.path.dirname(os.path.abspath(__file__))
    paths = [p for p in sys.path if os.path.abspath(p or os.getcwd()) != here]
    spec = importlib.machinery.PathFinder.find_spec(name, paths)
    if spec is None or spec.loader is None:
        return None
    module = importlib.util.module_from_spec(spec)
    saved = sys.modules.get(name)
    sys.modules[name] = module
    try:
        spec.loader.exec_module(module)
    except Exception:
        if saved is not None:
            sys.modules[name] = saved
        else:
            sys.modules.pop(name, None)
        return None
    return module
