"""tests-only stub: common/image_util.py imports imageio at module scope; nothing on the hot path calls it."""
def mimsave(*a, **k):
    raise NotImplementedError
