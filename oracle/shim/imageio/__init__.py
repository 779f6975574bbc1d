"""tests-only stub: common/image_util.py imports imageio at module scope; nothing on the hot path calls it."""
def mimsave(*a, **k):
    raise NotImplementedError


class _NullWriter:
    """get_writer(...) stand-in for the MP4 export of common/image_util.py:98-106 (the logger test only needs the call to succeed)."""

    def __init__(self, path, **kw):
        self.path, self.frames = path, 0

    def append_data(self, frame):
        self.frames += 1

    def close(self):
        pass


def get_writer(path, *a, **k):
    return _NullWriter(path, **k)
