"""Drop-in for the reference's `tdmpc2/tdmpc2.py` import (`from tdmpc2 import TDMPC2`,
reference evaluate.py:15): put this directory ahead of the reference on sys.path."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tdmpc2_b200.tdmpc2 import TDMPC2  # noqa: E402,F401
