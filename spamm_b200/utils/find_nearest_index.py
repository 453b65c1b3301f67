import numpy as np


def find_nearest_index(input_list, value):
    """Index of the entry closest to `value` (first one on ties); host-side
    helper used while resolving prior bounds (reference
    utils/find_nearest_index.py:21-34)."""
    arr = np.asarray(input_list, dtype=float)
    return int(np.argmin(np.abs(arr - value)))


def find_nearest(input_list, value):
    return input_list[find_nearest_index(input_list, value)]
