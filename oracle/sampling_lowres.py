"""numpy restatement of the reference's low-resolution farthest point sampling.

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Follows Preprocessing/preprocessing_sampling_lowres.py:14-42 line by line
(the reference functions are numba-jitted numpy code; float32 points, float64 running minimum).

Parity pin: tests/golden/ref_sampling_lowres.npz, produced by running the UNMODIFIED reference functions as plain
numpy (tests/golden/make_ref_sampling_lowres_golden.py; numba / h5py are not installed here and are stubbed at
import time -- ``numba.jit`` by the identity decorator).
"""
import numpy as np


def furthest_point_sampling(input_points, index_query_points1, nb_query_points):
    """:14-26."""
    num_points, _ = input_points.shape
    index_query_points2 = np.zeros(nb_query_points, dtype=np.int32)
    min_distances = 10 ** 6 * np.ones(num_points, dtype=np.float64)
    min_distances[index_query_points1] = 0
    index = np.argmax(min_distances)
    for i in range(nb_query_points):
        index_query_points2[i] = index
        additional_distances = np.sqrt(np.sum((input_points - input_points[index]) ** 2, axis=1))
        min_distances = np.minimum(min_distances, additional_distances)
        index = np.argmax(min_distances)
    return index_query_points2


def furthest_point_sampling_per_label(input_points, labels):
    """:28-42."""
    num_points, _ = input_points.shape
    unique_labels = np.unique(labels)
    index_query_points = np.zeros(len(unique_labels), dtype=np.int32)
    min_distances = 10 ** 6 * np.ones(num_points, dtype=np.float64)
    index = np.random.randint(0, num_points)
    for i in range(len(unique_labels)):
        label = labels[index]
        index_query_points[i] = index
        additional_distances = np.sqrt(np.sum((input_points - input_points[index]) ** 2, axis=1))
        min_distances = np.minimum(min_distances, additional_distances)
        min_distances[labels == label] = 0
        index = np.argmax(min_distances)
    return index_query_points
