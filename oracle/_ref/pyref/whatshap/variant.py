"""Stub: the dataclass whatshap.core imports as `.variant.Variant` (same three fields)."""
from whatshap_b200.variant import Variant  # noqa: F401
