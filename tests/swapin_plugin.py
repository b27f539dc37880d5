"""pytest plugin used by test_reference_suite_with_swapin.py: before the reference's test modules are
imported, replace `whatshap.core.PedigreeDPTable` / `GenotypeDPTable` by this repository's swap-in classes (the CPU checker
stands in for the per-call CUDA solve; the CUDA path itself is held to the same checker by the GPU tests)
and the host steps around the DP by this repository's: `whatshap.readselect.readselection`,
`whatshap.priorityqueue.PriorityQueue`, `whatshap.pedigree.find_recombination` / `centimorgen_to_phred`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.environ["WHMEC_PYREF"])

import whatshap.core as core  # noqa: E402
import whatshap.pedigree  # noqa: E402
import whatshap.priorityqueue  # noqa: E402
import whatshap.readselect  # noqa: E402  (Cython modules re-check core's extension types on import: before core is patched)

try:  # the compiled reference-side binding, if it was built (integration/build_bridge.py): before core is patched
    import whatshap_bridge  # noqa: E402,F401
except ImportError:
    pass

from oracle import checker  # noqa: E402
from whatshap_b200 import adapters  # noqa: E402

core.PedigreeDPTable = adapters.make_dp_table_class(core, solver=checker.port().solve)
# the genotyping DP: the kernels' per-cell code + host packer stepped on the host (tests/emul); the exact genotype
# priors reach the adapter through a recording Pedigree subclass (the real class has no getters)
import emul_genotype  # noqa: E402

core.Pedigree = adapters.recording_pedigree(core.Pedigree)
core.GenotypeDPTable = adapters.make_genotype_table_class(core, solver=lambda p: emul_genotype.genotype(p)[0])

from whatshap_b200 import pedigree as my_pedigree  # noqa: E402
from whatshap_b200 import priorityqueue as my_queue  # noqa: E402
from whatshap_b200 import readselect as my_select  # noqa: E402

whatshap.readselect.readselection = my_select.readselection
whatshap.priorityqueue.PriorityQueue = my_queue.PriorityQueue
whatshap.pedigree.find_recombination = my_pedigree.find_recombination
whatshap.pedigree.centimorgen_to_phred = my_pedigree.centimorgen_to_phred
whatshap.pedigree.RecombinationEvent = my_pedigree.RecombinationEvent
