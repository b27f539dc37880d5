"""SASS of the hottest loop of a kernel with the executed counts of an `ncu --set full --import-source on` report.
    python profiles/sass_hotloop.py gpurun_out/r02m_tile_cfg3.ncu-rep > profiles/r02m_tile_cfg3_sass_hotloop.txt
The hottest loop = the maximal run of consecutive instructions whose executed count equals the kernel's most frequent non-trivial
count (the steady-state column loop of tile_panel_kernel).  Prints the opcode mix by pipe next to the listing."""
import collections
import csv
import io
import subprocess
import sys

ALU = {"VIADDMNMX", "VIMNMX", "IADD3", "SHF", "LOP3", "PRMT", "SEL", "ISETP", "VIADD", "LEA", "PLOP3", "POPC", "FLO", "BREV", "MOV", "IABS"}
FMA = {"IMAD"}
LSU = {"LDS", "STS", "LDG", "STG", "LDC", "LDCU", "ATOMS", "ATOMG", "RED"}


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    execd = [int(r[ix["Instructions Executed"]]) for r in data]
    total = sum(execd)
    weight = collections.Counter()
    for e in execd:
        if e:
            weight[e] += e
    hot = weight.most_common(1)[0][0]
    # longest run of rows with that count
    best, cur, start = (0, 0), 0, 0
    for i, e in enumerate(execd + [0]):
        if e == hot:
            if cur == 0:
                start = i
            cur += 1
        else:
            if cur > best[0]:
                best = (cur, start)
            cur = 0
    n, s0 = best
    print("# %s" % rows[0][1][:160])
    print("# hottest loop: %d instructions, each executed %d times (warp level) = %.1f %% of the %d warp instructions of the launch"
          % (n, hot, 100.0 * n * hot / total, total))
    mix = collections.Counter()
    for r in data[s0:s0 + n]:
        toks = r[ix["Source"]].split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        mix[op] += 1
    pipe = collections.Counter()
    for op, c in mix.items():
        pipe["alu" if op in ALU else "fma" if op in FMA else "lsu" if op in LSU else "other"] += c
    print("# by pipe: " + ", ".join("%s %d" % kv for kv in pipe.most_common()))
    print("# by opcode: " + ", ".join("%s %d" % kv for kv in mix.most_common()))
    print("# columns: executed (warp level), stall samples, SASS")
    for r in data[s0:s0 + n]:
        print("%9s %5s   %s" % (r[ix["Instructions Executed"]], r[ix["# Samples"]], r[ix["Source"]].strip()))


if __name__ == "__main__":
    main(sys.argv[1])
