"""`Variant` record returned when indexing a `Read` (reference: whatshap/variant.py:4-10)."""
from dataclasses import dataclass


@dataclass
class Variant:
    """A single variant on a read"""

    position: int
    allele: int
    quality: int
