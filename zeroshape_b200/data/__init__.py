"""Input pipeline and on-disk formats either side of the hot path: mirrors of the reference's
demo.py:20-75 helpers and data/synthetic.py, with the pixel work on the device."""
