"""NumPy stand-in for the parts of ddsp v3.7.0 the reference's hot path uses.
Restated from the published sources (not vendored in /root/reference).
Test infrastructure."""
from . import core, processors, synths, effects  # noqa: F401
