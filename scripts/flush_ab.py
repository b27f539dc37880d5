"""A/B of WHMEC_FLUSH_UPLOAD (host threads write their part of the upload arrays back to memory before the DMA) on the end-to-end
time of whmec_solve, cfg3 and cfg5; three alternating rounds per setting.  GPU box."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from e2e_breakdown import CHILD  # noqa: E402

for name in ("cfg3", "cfg5"):
    for rnd in range(3):
        for flush in ("0", "1"):
            env = dict(os.environ, WHMEC_TIMING="1", WHMEC_FLUSH_UPLOAD=flush)
            r = subprocess.run([sys.executable, "-c", CHILD, name], env=env, capture_output=True, text=True, timeout=200)
            solve = [l for l in r.stderr.splitlines() if "solve:" in l][-1:]
            print(name, "flush", flush, (r.stdout.strip().splitlines() or ["?"])[-1], "|", solve[0].strip() if solve else "", flush=True)
