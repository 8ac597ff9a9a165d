"""usage: python tools/line_stalls.py <ncu --page source --csv export> [mangled kernel name substring, default residual_kernel<5,2,0,1>]
Aggregates warp-stall samples by CUDA source line: the SASS order of the profile is matched with `nvdisasm -g` of the
in-tree library (same build), whose `//## File ..., line N` markers give each instruction its line."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNEL = "_ZN2sg15residual_kernelILi5ELi2ELi0ELb1EEEvNS_9ResParamsE"


def line_table():
    global KERNEL
    if len(sys.argv) > 2:
        KERNEL = sys.argv[2]
    with tempfile.TemporaryDirectory() as wd:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "structured_b200", "libstructured_gpu.so")], cwd=wd, check=True, stdout=subprocess.DEVNULL)
        cubin = [f for f in os.listdir(wd) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "-g", os.path.join(wd, cubin)], capture_output=True, text=True).stdout
    seq, cur, on = [], None, False
    for l in dis.split("\n"):
        if l.startswith("\t.section\t.text."):
            on = KERNEL in l
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
        elif re.search(r'/\*[0-9a-f]{4,5}\*/\s+[@A-Z]', l):
            seq.append(cur)
    return seq


def main():
    seq = line_table()
    rows = list(csv.reader(open(sys.argv[1])))
    hdr = rows[1]
    data = [r for r in rows[2:] if len(r) >= len(hdr)]
    idx = {h: i for i, h in enumerate(hdr)}
    assert len(seq) == len(data), (len(seq), len(data), "profile and library are different builds")
    agg = collections.defaultdict(collections.Counter)
    tot = 0
    for k, r in enumerate(data):
        n = int(r[idx["# Samples"]])
        tot += n
        agg[seq[k]]["n"] += n
        agg[seq[k]]["inst"] += 1
        agg[seq[k]]["exec"] += int(r[idx["Instructions Executed"]])
        for h in hdr:
            if h.startswith("stall_") and "Not Issued" not in h:
                agg[seq[k]][h[6:]] += int(r[idx[h]])
    texec = sum(c["exec"] for c in agg.values())
    print("total samples", tot, "warp instructions executed", texec)
    for key, c in sorted(agg.items(), key=lambda kv: -kv[1]["n"])[:int(os.environ.get("TOP", "45"))]:
        print(key, "n", c["n"], round(100 * c["n"] / tot, 1), "inst", c["inst"], "exec%", round(100 * c["exec"] / max(texec, 1), 1),
              {k: v for k, v in c.items() if k not in ("n", "inst", "exec") and v > 0.08 * c["n"]})


if __name__ == "__main__":
    main()
