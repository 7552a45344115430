"""Test-infrastructure stub (oracle only): plotting is never exercised."""
rcParams = {}


def use(*args, **kwargs):
    pass
