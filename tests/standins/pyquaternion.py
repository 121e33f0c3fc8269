"""Stand-in: wild_completion/opt_visualizer.py imports Quaternion at module top; only the interactive GUI uses it."""


class Quaternion:
    def __init__(self, *a, **k):
        raise NotImplementedError("pyquaternion stand-in: the interactive visualiser is not available in the tests")
