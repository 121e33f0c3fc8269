class Dict(dict):
    """Stand-in for addict.Dict (only subclassed by wild_completion/utils.py:524 ForceKeyErrorDict)."""

    def __getattr__(self, k):
        return self[k]

    __setattr__ = dict.__setitem__
