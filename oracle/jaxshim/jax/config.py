"""`from jax.config import config; config.update("jax_enable_x64", True)` (admp/settings.py:3-9):
the shim is float64 throughout, so this only records the request."""


class _Config:
    def __init__(self):
        self.values = {}

    def update(self, key, value):
        self.values[key] = value


config = _Config()


def update(key, value):
    config.update(key, value)
