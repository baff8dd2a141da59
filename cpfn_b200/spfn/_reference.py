"""Everything of the reference's ``SPFN.losses_implementation`` / ``SPFN.metric_implementation`` that is NOT on
the hot path (Hungarian matching, mIoU / normal / type losses, loss collection, metrics ...) stays the
reference's own code: when the reference checkout is on ``sys.path`` its module file is loaded privately, the
hot-path functions in it are replaced by this package's, and attribute look-ups that this package does not
define are forwarded to it (PEP 562 ``__getattr__`` of the two mirror modules).  Without the checkout those
names raise AttributeError -- this package does not restate them."""
import importlib.util
import os
import sys

_loaded = {}
_HERE = os.path.dirname(os.path.abspath(__file__))


def find(name):
    for root in sys.path:
        path = os.path.join(root or ".", "SPFN", name + ".py")
        if os.path.isfile(path) and os.path.dirname(os.path.abspath(path)) != _HERE:
            return path
    return None


def load(name, overrides):
    """The reference's SPFN/<name>.py as a private module, with ``overrides`` (dict name -> object) patched in."""
    if name in _loaded:
        return _loaded[name]
    path = find(name)
    if path is None:                      # not cached: the checkout may be put on sys.path later
        return None
    spec = importlib.util.spec_from_file_location("cpfn_b200.spfn._ref_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)          # its `from SPFN import ...` lines resolve to whatever SPFN is installed
    for k, v in overrides.items():
        setattr(mod, k, v)
    _loaded[name] = mod
    return mod


def forwarder(module_globals, name, hot_names):
    """PEP 562 ``__getattr__`` for the mirror module ``cpfn_b200.spfn.<name>``: names this package does not define
    (``create_primitive_from_dict``, ``extract_parameter_data_as_dict``, ``compute_parameter_loss``, the TensorFlow
    twins ...: host-side metadata and loss glue outside the hot path) are looked up in the reference's own
    ``SPFN/<name>.py``, loaded privately with this package's ``hot_names`` patched into it."""
    def __getattr__(attr):
        ref = load(name, {k: module_globals[k] for k in hot_names})
        if ref is not None and hasattr(ref, attr):
            return getattr(ref, attr)
        raise AttributeError("cpfn_b200.spfn.%s has no '%s' (not a hot-path function; put the reference checkout "
                             "on sys.path to use the reference's own)" % (name, attr))
    return __getattr__
