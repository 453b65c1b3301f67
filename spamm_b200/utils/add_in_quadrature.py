import numpy as np


def add_in_quadrature(data_in):
    """sqrt(sum_i x_i^2) over a sequence of arrays (utils/add_in_quadrature.py:5)."""
    total = 0.
    for data in data_in:
        total = total + np.array(data) ** 2
    return np.sqrt(total)
