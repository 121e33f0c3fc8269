"""Stand-in: wild_completion/utils.py imports plyfile at module top; the hot path never calls it."""
