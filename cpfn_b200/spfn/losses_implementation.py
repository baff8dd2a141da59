"""``compute_parameters(P, W, X, classes)`` with the dictionary keys of the reference
(SPFN/losses_implementation.py:255-278).  All requested classes come out of ONE fused
kernel sequence (cpfn_fit_primitives) instead of four serial fitters that each re-tile P / W / X."""
from . import fit

_CLASS_KEYS = {
    "plane": ("plane_normal", "plane_center"),
    "sphere": ("sphere_center", "sphere_radius_squared"),
    "cylinder": ("cylinder_axis", "cylinder_center", "cylinder_radius_squared"),
    "cone": ("cone_apex", "cone_axis", "cone_half_angle"),
}


def compute_parameters(P, W, X, classes=['plane', 'sphere', 'cylinder', 'cone']):
    for class_ in classes:
        if class_ not in _CLASS_KEYS:
            raise NotImplementedError
    return compute_parameters_packed(P, W, X, classes)[0]


def compute_parameters_packed(P, W, X, classes=['plane', 'sphere', 'cylinder', 'cone']):
    """(parameters dict, the one packed buffer behind it | None) -- see fit.fit_primitives_packed."""
    r, packed = fit.fit_primitives_packed(P, W, X)
    parameters = {}
    for class_ in classes:
        for key in _CLASS_KEYS[class_]:
            parameters[key] = r[key]
    return parameters, packed
