import numpy as np


def runningMeanFast(x, N):
    """Boxcar running mean with the reference's edge convention
    (utils/runningmeanfast.py:19): full convolution, first N-1 samples dropped."""
    box = np.ones((N,)) / N
    return np.convolve(x, box)[(N - 1):]
