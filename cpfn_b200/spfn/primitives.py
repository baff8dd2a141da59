"""``SPFN.primitives`` (numpy value objects describing ground-truth primitives: host-side metadata, not on the hot
path): the reference's own file, loaded from its checkout on sys.path (see _reference.py).  Exists so that the
reference's fitter modules, whose non-hot-path helpers are forwarded to, can ``from SPFN.primitives import Plane``
while ``SPFN`` is this package (cpfn_b200.dropin, level "full")."""
from . import _reference

_ref = _reference.load("primitives", {})
if _ref is None:
    raise ImportError("cpfn_b200.spfn.primitives forwards to the reference's SPFN/primitives.py: put the reference "
                      "checkout on sys.path")
globals().update({k: v for k, v in vars(_ref).items() if not k.startswith("__")})
