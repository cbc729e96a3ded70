"""Fall-through of the mirror's modules to the reference tree (creste_public_b200.install_as_creste(reference_root)).

A mirror module implements the hot-path part of its reference namesake (e.g. creste/utils/train_utils.py has the FOV
mask and dict helpers, not the dataset samplers).  When the mirror is overlaid on a reference checkout, any name the
mirror module does not define is looked up in the reference's file of the same relative path, loaded once under a
private module name -- so `from creste.utils.utils import make_labels_contiguous_vectorized` keeps working for the
reference's own scripts while `remap_labels_in_batch` comes from the mirror."""
import importlib.util
import os
import sys

REFERENCE_ROOT = None
_loaded = {}


def _reference_module(relpath):
    if REFERENCE_ROOT is None:
        return None
    if relpath not in _loaded:
        path = os.path.join(REFERENCE_ROOT, "creste", relpath)
        if not os.path.isfile(path):
            _loaded[relpath] = None
        else:
            name = "creste._reference." + relpath[:-3].replace("/", ".")
            spec = importlib.util.spec_from_file_location(name, path)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            _loaded[relpath] = mod           # registered before exec: import cycles resolve to the partial module
            spec.loader.exec_module(mod)
    return _loaded[relpath]


def fallback(module_name, relpath):
    """-> a module-level __getattr__ (PEP 562) for the mirror module `module_name`."""
    def __getattr__(name):
        if name.startswith("__"):
            raise AttributeError(name)
        ref = _reference_module(relpath)
        if ref is not None and hasattr(ref, name):
            return getattr(ref, name)
        raise AttributeError(f"module {module_name!r} has no attribute {name!r}"
                             + ("" if REFERENCE_ROOT else " (the mirror is not overlaid on a reference checkout: "
                                                          "creste_public_b200.install_as_creste(reference_root))"))
    return __getattr__
