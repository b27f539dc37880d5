"""`PriorityQueue` with the interface and tie behaviour of the reference's
`whatshap.priorityqueue.PriorityQueue` (whatshap/priorityqueue.pyx): a positional binary max-heap
over (score, item), scores being ints or sequences of ints compared lexicographically (a proper prefix
is smaller), items ints that can be re-scored in place.

Which of several equal scores is popped first depends on the heap's sift rules, so they are part of
the contract: a child moves above its parent only if strictly larger; of two children the right one is
preferred only if strictly larger than the left.  The read selection (`readselect.py`) runs the same
rules natively (csrc/readselect.cpp); this class is the general-purpose Python form and the yardstick
the native heap is tested against."""
from __future__ import annotations

from typing import Dict, List, Tuple


def _as_score(score) -> Tuple[int, ...]:
    if isinstance(score, int) and not isinstance(score, bool):
        return (score,)
    try:
        result = tuple(score)
    except TypeError:
        result = None
    if result is None or not all(isinstance(x, int) and not isinstance(x, bool) for x in result):
        raise ValueError("Score parameter must be either int, or an iterable object yielding ints")
    return result


class PriorityQueue:
    def __init__(self):
        self._scores: List[Tuple[int, ...]] = []
        self._items: List[int] = []
        self._slot: Dict[int, int] = {}

    # tuple comparison is exactly the reference's rule (element-wise, then the shorter one is lower)
    def _lower(self, a: int, b: int) -> bool:
        return self._scores[a] < self._scores[b]

    def _swap(self, a: int, b: int) -> None:
        self._scores[a], self._scores[b] = self._scores[b], self._scores[a]
        self._items[a], self._items[b] = self._items[b], self._items[a]
        self._slot[self._items[a]] = a
        self._slot[self._items[b]] = b

    def _up(self, i: int) -> None:
        while i > 0:
            parent = (i - 1) // 2
            if not self._lower(parent, i):
                return
            self._swap(parent, i)
            i = parent

    def _down(self, i: int) -> None:
        n = len(self._items)
        while True:
            left, right = 2 * i + 1, 2 * i + 2
            if right < n:
                child = right if self._lower(left, right) else left
            elif left < n:
                child = left
            else:
                return
            if not self._lower(i, child):
                return
            self._swap(child, i)
            i = child

    def push(self, score, item: int) -> None:
        """Add `item` with `score`."""
        self._scores.append(_as_score(score))
        self._items.append(int(item))
        self._slot[int(item)] = len(self._items) - 1
        self._up(len(self._items) - 1)

    def pop(self):
        """Remove the entry with the largest score; returns (score, item), the score as an int if it has
        one component, else as a tuple."""
        if not self._items:
            raise IndexError("PriorityQueue empty.")
        score, item = self._scores[0], self._items[0]
        last_score, last_item = self._scores.pop(), self._items.pop()
        del self._slot[item]
        if self._items:
            self._scores[0], self._items[0] = last_score, last_item
            self._slot[last_item] = 0
            self._down(0)
        return (score[0] if len(score) == 1 else score), item

    def change_score(self, item: int, new_score) -> None:
        new = _as_score(new_score)
        i = self._slot[int(item)]
        old, self._scores[i] = self._scores[i], new
        if old < new:
            self._up(i)
        else:
            self._down(i)

    def get_score_by_item(self, item):
        """Score of `item`, or None if it is not queued."""
        i = self._slot.get(item)
        if i is None:
            return None
        score = self._scores[i]
        return score[0] if len(score) == 1 else score

    def __len__(self) -> int:
        return len(self._items)

    def is_empty(self) -> bool:
        return not self._items
