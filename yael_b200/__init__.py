"""yael_b200 -- B200-native (sm_100a) implementation of yael's distance / top-k / k-means /
Hamming hot path behind yael's own C API.  See DESIGN.md.

    from yael_b200 import ynumpy          # numpy front-end, same call shapes as yael.ynumpy
    from yael_b200 import dist            # one-process-per-GPU sharded kNN / k-means
"""
from ._lib import LIB_PATH, YaelB200Error, build, lib  # noqa: F401

__all__ = ["LIB_PATH", "YaelB200Error", "build", "lib"]
