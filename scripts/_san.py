import sys
sys.path.insert(0, ".")
from oracle import checker
from whatshap_b200 import _lib, synth
ck = checker.best()
for prob in (synth.sliding_window(60, 17, block_len=60, seed=3), synth.sliding_window(48, 20, block_len=48, seed=5),
             synth.sliding_window(90, 16, block_len=45, seed=3, gap=0.1, max_phred=3), synth.sliding_window(300, 12, block_len=100, seed=1),
             synth.trio(60, 3, block_len=30, seed=3)):
    got, st = _lib.solve(prob)
    assert got.same_as(ck.solve(prob))
    print("ok", st["path_kind"], st["max_active"])
