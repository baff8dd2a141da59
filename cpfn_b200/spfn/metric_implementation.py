"""The residue-based metrics of the reference's ``SPFN/metric_implementation.py`` on the fused residue kernels:
``get_residual_loss`` (:76-81) and ``compute_P_coverage`` (:409-415).  Other names are forwarded to the
reference's own file when its checkout is on sys.path (_reference.py)."""
import torch

from . import _reference, losses_implementation, residues, seg


_LUT = {}


def sqrt_safe(x):
    return torch.sqrt(torch.abs(x) + 1e-10)


def get_residual_loss(parameters, matching_indices, points_per_instance, T, classes=['plane', 'sphere', 'cylinder', 'cone']):
    """:76-81.  T int64 [B,K] = type index (into ``classes``) of every primitive.  -> [B,K,N']"""
    batch_size, num_primitives, num_primitive_points, _ = points_per_instance.shape
    _, residue_per_point_array = losses_implementation.compute_residue_loss(
        parameters, matching_indices, points_per_instance, torch.gather(T, 1, matching_indices), classes=classes)
    residue_per_point_array = torch.gather(
        residue_per_point_array, 3,
        T.view(batch_size, num_primitives, 1, 1).expand(batch_size, num_primitives, num_primitive_points, 1)).squeeze(3)
    return sqrt_safe(residue_per_point_array)


def compute_P_coverage(P, T, matching_indices, predicted_parameters, epsilon, classes=['plane', 'sphere', 'cylinder', 'cone']):
    """:409-415.  P [B,N,3], T int64 [B,K].  ``epsilon`` may be one float (-> [B], as the reference) or a list of up
    to four (-> [B, n]); the per-point minimum over the primitives is taken inside the kernel."""
    prim_type = torch.gather(T, 1, matching_indices)                      # the reference passes this as ``T`` (:412)
    key = (tuple(classes), P.device)
    lut = _LUT.get(key)
    if lut is None:
        lut = _LUT[key] = torch.tensor([residues.CLASS_ID[c] for c in classes], dtype=torch.int64, device=P.device)
    single = not isinstance(epsilon, (list, tuple))
    cov = residues.p_coverage(P, lut[prim_type], matching_indices, predicted_parameters, [epsilon] if single else epsilon)
    return cov[:, 0] if single else cov

def hungarian_matching(W_pred, I_gt):
    """SPFN/metric_implementation.py:9-30 on the device: (matching_indices int64 [B,K], mask bool [B,K])."""
    return seg.hungarian_matching(W_pred, I_gt, with_mask=True)


def __getattr__(name):
    ref = _reference.load("metric_implementation", {"get_residual_loss": get_residual_loss,
                                                    "compute_P_coverage": compute_P_coverage,
                                                    "hungarian_matching": hungarian_matching})
    if ref is not None and hasattr(ref, name):
        return getattr(ref, name)
    raise AttributeError("cpfn_b200.spfn.metric_implementation has no '%s' (not a hot-path function; put the "
                         "reference checkout on sys.path to use the reference's own)" % name)
