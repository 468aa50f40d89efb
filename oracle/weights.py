"""TEST INFRASTRUCTURE (oracle) -- the seeded synthetic weights / inputs live in the package
(``athtd_b200.synthetic``, one definition for bench, smoke and tests); re-exported here so oracle-side
scripts keep reading ``oracle.weights``."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
import athtd_b200  # noqa: E402,F401  (import alias of the package directory)
from athtd_b200.synthetic import (CH, live_param_table, make_inputs, make_state_dict,  # noqa: E402,F401
                                  state_dict_checksum)
