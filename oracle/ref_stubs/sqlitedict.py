"""Test-infrastructure stub (oracle only): in-memory stand-in for on-disk shelves,
which the hot path (in-memory Empirical, OnlineDataset) never touches."""


class SqliteDict(dict):
    def __init__(self, filename=None, *args, **kwargs):
        super().__init__()
        self.filename = filename

    def sync(self):
        pass

    def commit(self, *args, **kwargs):
        pass

    def close(self, *args, **kwargs):
        pass
