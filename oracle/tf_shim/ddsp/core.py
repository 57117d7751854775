"""ddsp.core stand-in = the oracle's restatement."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(
    os.path.abspath(__file__))))))
from oracle.ddsp_core_np import *  # noqa: F401,F403,E402
from oracle.ddsp_core_np import nested_lookup  # noqa: F401,E402
