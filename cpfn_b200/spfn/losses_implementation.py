"""``compute_parameters(P, W, X, classes)`` with the dictionary keys of the reference
(SPFN/losses_implementation.py:255-278).  All requested classes come out of ONE fused
kernel sequence (cpfn_fit_primitives) instead of four serial fitters that each re-tile P / W / X.
``compute_residue_loss`` (:351-387) evaluates the residues of all requested types in one kernel
(cpfn_primitive_residues) when no gradient is needed, and with the reference's element-wise formulas under
autograd.  Any other name of the reference module is forwarded to the reference's own file when its checkout is
on sys.path (see _reference.py): those functions are not on the hot path."""
import torch

from . import _reference, fit, residues, seg
from . import plane_fitter, sphere_fitter, cylinder_fitter, cone_fitter

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


def _gather3(t, matching_indices):
    b, k = matching_indices.shape
    return torch.gather(t, 1, matching_indices.unsqueeze(2).expand(b, k, 3)).unsqueeze(2)


def _gather1(t, matching_indices):
    return torch.gather(t, 1, matching_indices).unsqueeze(2)


def compute_residue_loss(parameters, matching_indices, points_per_instance, T_gt,
                         classes=['plane', 'sphere', 'cylinder', 'cone']):
    """SPFN/losses_implementation.py:351-387.  parameters: dict of [B,K,(3)]; matching_indices int64 [B,K];
    points_per_instance [B,K,N',3]; T_gt int64 [B,K] (index into ``classes``).
    Returns (residue_loss [B,K], residue_per_point_array [B,K,N',T])."""
    for class_ in classes:
        if class_ not in _CLASS_KEYS:
            raise NotImplementedError
    needs_grad = torch.is_grad_enabled() and (points_per_instance.requires_grad or any(
        parameters[key].requires_grad for class_ in classes for key in _CLASS_KEYS[class_]))
    if not needs_grad:
        residue_per_point_array, residue_losses = residues.residues(parameters, matching_indices, points_per_instance,
                                                                    classes)
    else:
        m, per_class = matching_indices, []
        for class_ in classes:
            if class_ == 'plane':
                r = plane_fitter.compute_residue_single(_gather3(parameters['plane_normal'], m),
                                                        _gather1(parameters['plane_center'], m), points_per_instance)
            elif class_ == 'sphere':
                r = sphere_fitter.compute_residue_single(_gather3(parameters['sphere_center'], m),
                                                         _gather1(parameters['sphere_radius_squared'], m),
                                                         points_per_instance)
            elif class_ == 'cylinder':
                r = cylinder_fitter.compute_residue_single(_gather3(parameters['cylinder_axis'], m),
                                                           _gather3(parameters['cylinder_center'], m),
                                                           _gather1(parameters['cylinder_radius_squared'], m),
                                                           points_per_instance)
            else:
                r = cone_fitter.compute_residue_single(_gather3(parameters['cone_apex'], m),
                                                       _gather3(parameters['cone_axis'], m),
                                                       _gather1(parameters['cone_half_angle'], m), points_per_instance)
            per_class.append(r)
        residue_per_point_array = torch.stack(per_class, dim=3)
        residue_losses = torch.stack([torch.mean(r, dim=2) for r in per_class], dim=2)
    residue_loss = torch.gather(residue_losses, 2, T_gt.unsqueeze(2)).squeeze(2)
    return residue_loss, residue_per_point_array

def hungarian_matching(W_pred, I_gt):
    """SPFN/losses_implementation.py:11-30 on the device (spfn/seg.py): no host round trip, no scipy."""
    return seg.hungarian_matching(W_pred, I_gt)


def compute_miou_loss(W, I_gt, matching_indices, div_eps=1e-10):
    """SPFN/losses_implementation.py:77-89 (spfn/seg.py)."""
    return seg.compute_miou_loss(W, I_gt, matching_indices, div_eps)


def __getattr__(name):
    ref = _reference.load("losses_implementation", {"compute_parameters": compute_parameters,
                                                    "compute_residue_loss": compute_residue_loss,
                                                    "hungarian_matching": hungarian_matching,
                                                    "compute_miou_loss": compute_miou_loss})
    if ref is not None and hasattr(ref, name):
        return getattr(ref, name)
    raise AttributeError("cpfn_b200.spfn.losses_implementation has no '%s' (not a hot-path function; put the "
                         "reference checkout on sys.path to use the reference's own)" % name)
