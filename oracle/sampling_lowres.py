"""numpy restatement of the reference's low-resolution farthest point sampling.

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Same arithmetic and the same decisions as the two numba functions of
Preprocessing/preprocessing_sampling_lowres.py (cited per step below), written around one shared sweep:
float32 points, a float64 running minimum that starts at 1e6, true distances sqrt(sum((p - q)^2)) in the points'
dtype, np.argmax (first maximum) for the pick.

Parity pin: tests/golden/ref_sampling_lowres.npz, produced by running the UNMODIFIED reference functions as plain
numpy (tests/golden/make_ref_sampling_lowres_golden.py; numba / h5py are not installed here and are stubbed at
import time -- ``numba.jit`` by the identity decorator).
"""
import numpy as np

_FAR = 10 ** 6


def _fold(points, running_min, picked):
    """One sweep: distances of every point to ``picked`` folded into the running minimum (:22-23, :38-39)."""
    gap = points - points[picked]
    return np.minimum(running_min, np.sqrt(np.sum(gap ** 2, axis=1)))


def furthest_point_sampling(input_points, index_query_points1, nb_query_points):
    """:14-26 -- the seeds only START at distance 0 (:18); their neighbourhoods are not excluded."""
    running_min = np.full(len(input_points), _FAR, dtype=np.float64)
    running_min[index_query_points1] = 0
    chosen = np.zeros(nb_query_points, dtype=np.int32)
    for slot in range(nb_query_points):
        chosen[slot] = np.argmax(running_min)            # :19 / :24, recorded at :21
        running_min = _fold(input_points, running_min, chosen[slot])
    return chosen


def furthest_point_sampling_per_label(input_points, labels):
    """:28-42 -- one sample per distinct label; after a pick every point of that label drops out (:40)."""
    n_labels = len(np.unique(labels))
    running_min = np.full(len(input_points), _FAR, dtype=np.float64)
    chosen = np.zeros(n_labels, dtype=np.int32)
    pick = np.random.randint(0, len(input_points))       # :33
    for slot in range(n_labels):
        chosen[slot] = pick
        running_min = _fold(input_points, running_min, pick)
        running_min[labels == labels[pick]] = 0
        pick = np.argmax(running_min)                    # :41
    return chosen
