"""Test-infrastructure stub (oracle only): plotting is never exercised."""


def __getattr__(name):
    def _unavailable(*args, **kwargs):
        raise RuntimeError('matplotlib is stubbed in the oracle harness: ' + name)
    return _unavailable
