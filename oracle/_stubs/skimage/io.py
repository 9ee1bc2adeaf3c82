"""Stub: the label/gray arithmetic the goldens pin never reads image files."""


def imread(*a, **k):
    raise RuntimeError("skimage stub: no image IO in the oracle")


imread_collection = imread
