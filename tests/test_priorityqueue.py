"""`PriorityQueue` (whatshap_b200/priorityqueue.py) against the reference's known answers
(tests/test_priorityqueue.py) and against operation traces recorded from the reference's queue
(tests/golden/make_readselect_golden.py): every pop must return the same (score, item), ties included."""
import os

import numpy as np
import pytest

from whatshap_b200.priorityqueue import PriorityQueue

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "readselect.npz")


def drain(pq):
    out = []
    while not pq.is_empty():
        out.append(pq.pop())
    return out


def test_pops_come_out_by_score():
    pq = PriorityQueue()
    for score, item in ((10, "a"), (5, "b"), (12, "c"), (3, "d")):
        pq.push(score, ord(item))
    assert len(pq) == 4 and pq.get_score_by_item(ord("d")) == 3 and pq.get_score_by_item(ord("z")) is None
    assert drain(pq) == [(12, ord("c")), (10, ord("a")), (5, ord("b")), (3, ord("d"))]
    with pytest.raises(IndexError):
        pq.pop()


def test_change_score_moves_entries_both_ways():
    pq = PriorityQueue()
    pq.push(10, 1)
    pq.push(5, 2)
    pq.change_score(1, 2)
    pq.push(12, 3)
    pq.push(3, 4)
    pq.change_score(3, 1)
    pq.change_score(4, 15)
    assert drain(pq) == [(15, 4), (5, 2), (2, 1), (1, 3)]
    for score, item in ((50, 1), (40, 2), (30, 3), (20, 4), (10, 5)):
        pq.push(score, item)
    pq.change_score(5, 100)
    pq.change_score(2, 45)
    assert pq.pop() == (100, 5)
    pq.push(60, 8)
    assert pq.pop() == (60, 8)
    pq.change_score(2, 40)
    assert pq.pop() == (50, 1) and pq.pop() == (40, 2)


def test_tuple_scores_compare_lexicographically():
    pq = PriorityQueue()
    for score, item in (((10, 0, 0), 1), ((10, 2, 6), 2), ((10, 3, 2), 3), ((10, 4, 3), 4), ((10, 2, 2), 5), ((10, 0, 2), 6)):
        pq.push(score, item)
    assert [pq.pop() for _ in range(5)] == [((10, 4, 3), 4), ((10, 3, 2), 3), ((10, 2, 6), 2), ((10, 2, 2), 5), ((10, 0, 2), 6)]
    pq.push((1, 10, 4), 7)
    pq.push((5, 0, 6), 8)
    pq.push((1, 8, 2), 9)
    pq.change_score(8, (100, 100, 100))
    pq.change_score(9, (0, 0, 0))
    assert pq.get_score_by_item(7) == (1, 10, 4)
    assert drain(pq) == [((100, 100, 100), 8), ((10, 0, 0), 1), ((1, 10, 4), 7), ((0, 0, 0), 9)]
    assert (1,) < (1, 0) and pq.is_empty()  # a proper prefix is the lower score (priorityqueue.pyx:18-21)
    with pytest.raises(ValueError):
        pq.push("high", 1)
    with pytest.raises(ValueError):
        pq.push((1, 2.5), 1)


def test_recorded_traces_of_the_reference_queue():
    """40 random traces (pushes, pops, re-scorings, look-ups; scores from a tiny range so that ties abound)."""
    z = np.load(GOLDEN)
    for t in range(int(z["pq.n"])):
        width = int(z["pq.width"][z["pq.width.off"][t]])
        ops = z["pq.ops"][z["pq.ops.off"][t]:z["pq.ops.off"][t + 1]].reshape(-1, 5)
        answers = iter(z["pq.answers"][z["pq.answers.off"][t]:z["pq.answers.off"][t + 1]].reshape(-1, 4).tolist())
        as_score = lambda s: tuple(s[:width]) if width > 1 else s[0]
        pq = PriorityQueue()
        for kind, item, *score in ops.tolist():
            if kind == 0:
                pq.push(as_score(score), item)
            elif kind == 1:
                got_score, got_item = pq.pop()
                want = next(answers)
                assert (got_item, got_score) == (want[0], as_score(want[1:])), t
            elif kind == 2:
                pq.change_score(item, as_score(score))
            else:
                want = next(answers)
                assert pq.get_score_by_item(item) == as_score(want[1:]), t
        assert pq.is_empty() and next(answers, None) is None
