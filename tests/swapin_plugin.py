"""pytest plugin used by test_reference_suite_with_swapin.py: before the reference's test modules are
imported, replace `whatshap.core.PedigreeDPTable` by this repository's swap-in class (the CPU checker
stands in for the per-call CUDA solve; the CUDA path itself is held to the same checker by the GPU tests)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.environ["WHMEC_PYREF"])

import whatshap.core as core  # noqa: E402

try:  # the compiled reference-side binding, if it was built (integration/build_bridge.py): before core is patched
    import whatshap_bridge  # noqa: E402,F401
except ImportError:
    pass

from oracle import checker  # noqa: E402
from whatshap_b200 import adapters  # noqa: E402

core.PedigreeDPTable = adapters.make_dp_table_class(core, solver=checker.port().solve)
