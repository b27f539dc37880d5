import sys
sys.path.insert(0, ".")
from oracle import checker
from whatshap_b200 import _lib, synth
ck = checker.best()
for prob in (synth.trio(60, 3, block_len=30, seed=3), synth.trio(40, 5, block_len=20, seed=4), synth.trio(24, 2, block_len=6, seed=5)):
    got, st = _lib.solve(prob)
    assert got.same_as(ck.solve(prob))
    print("ok", st["path_kind"], st["max_active"], st["kernel_launches"])
