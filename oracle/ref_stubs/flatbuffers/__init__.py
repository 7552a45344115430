"""Test-infrastructure stub (oracle only): PPX wire format is out of scope."""
from . import compat  # noqa: F401


class Builder:
    def __init__(self, *args, **kwargs):
        raise RuntimeError('flatbuffers is stubbed in the oracle harness')
