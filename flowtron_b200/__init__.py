"""flowtron_b200 — B200-native (sm_100a) implementation of the Flowtron AR-flow hot path."""
from . import _lib  # noqa: F401
