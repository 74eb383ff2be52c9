import numpy as np


def skew(vec):
    """[v]x  (reference image_processing/utils.py:3-8)."""
    x, y, z = vec
    return np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]])


def select(data, selectors):
    """Keep data[i] where selectors[i] is truthy (reference image_processing/utils.py:10-11)."""
    return [d for d, s in zip(data, selectors) if s]
