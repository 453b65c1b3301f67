import numpy as np


def gaussian(minimum, maximum, num=None):
    """Normal draw centred on the interval, sigma = 34.1 % of the half width
    (utils/draw_from_sample.py:5-11)."""
    mu = minimum + ((maximum - minimum) / 2.)
    sigma = (maximum - mu) * .341
    return np.random.normal(mu, sigma, num)
