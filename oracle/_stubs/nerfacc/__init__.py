"""Placeholder: nerfacc is not installed; its arithmetic is restated in oracle/ref_nerfacc.py."""


class PropNetEstimator:
    def __init__(self, *a, **k):
        pass
