"""Abstract contract of a phasing algorithm (reference: whatshap/types.py:7-15)."""
from abc import ABC, abstractmethod
from typing import List, Optional, Tuple


class PhasingAlgorithm(ABC):
    @abstractmethod
    def get_super_reads(self) -> Tuple[List["ReadSet"], Optional[List[int]]]: ...

    @abstractmethod
    def get_optimal_cost(self) -> int: ...

    @abstractmethod
    def get_optimal_partitioning(self) -> List[int]: ...
