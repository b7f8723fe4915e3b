"""Inert stand-in for `gym` (only what multiagent/environment.py touches)."""
from . import spaces
from .spaces import Space


class Env(object):
    metadata = {}

    def close(self):
        return None
