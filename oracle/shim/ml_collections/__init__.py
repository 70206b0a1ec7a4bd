"""Attribute-dict stand-in for ml_collections.ConfigDict (configs/*.py of the reference).
TEST INFRASTRUCTURE ONLY."""


class ConfigDict(dict):
    def __init__(self, initial=None):
        super().__init__()
        if initial:
            for k, v in initial.items():
                self[k] = ConfigDict(v) if isinstance(v, dict) and not isinstance(v, ConfigDict) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def lock(self):
        return self
