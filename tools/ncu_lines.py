#!/usr/bin/env python
"""Per-source-line instruction and stall-sample totals for one kernel of an .ncu-rep.

    python tools/ncu_lines.py REPORT.ncu-rep INVOCATION MANGLED_SUBSTRING [host.o [DEMANGLED_SUBSTRING]]

Joins `ncu --page source --csv` (SASS rows: executed instructions, stall samples) with
`nvdisasm -g` line info of the same function (built with -lineinfo), by instruction offset.
"""
import csv
import io
import re
import subprocess
import sys
import tempfile
from collections import defaultdict
from pathlib import Path


def sass_lines(obj: Path, needle: str):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", str(obj.resolve())], cwd=tmp, check=True, capture_output=True)
    cubin = next(Path(tmp).glob("*.cubin"))
    txt = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True, text=True).stdout
    out, cur, active = {}, None, False
    for line in txt.splitlines():
        if line.startswith("//---") and ".text." in line:
            active = needle in line
            cur = None
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', line)
        if m:
            cur = f"{Path(m.group(1)).name}:{m.group(2)}"
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            out[int(m.group(1), 16)] = (cur, m.group(2))
    return out


def main():
    rep, kid, needle = sys.argv[1], sys.argv[2], sys.argv[3]
    obj = Path(sys.argv[4] if len(sys.argv) > 4 else "polars-strsim_b200/csrc/host.o")
    lines = sass_lines(obj, needle)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    name_sub = sys.argv[5] if len(sys.argv) > 5 else ""
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    pick = next(i for i in starts if name_sub in rows[i][1])
    end = next((j for j in starts if j > pick), len(rows))
    rows = rows[pick:end]
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    ix = {h: i for i, h in enumerate(hdr)}
    body = rows[hdr_i + 1:]
    base = int(body[0][0], 16)
    inst = defaultdict(int)
    thr = defaultdict(int)
    samp = defaultdict(int)
    stall_cols = [h for h in hdr if h.startswith("stall_")]
    stalls = defaultdict(lambda: defaultdict(int))
    total_inst = total_samp = 0
    for r in body:
        off = int(r[0], 16) - base
        where = lines.get(off, (None, ""))[0] or "?"
        n = int(r[ix["Instructions Executed"]] or 0)
        s = int(r[ix["# Samples"]] or 0)
        inst[where] += n
        thr[where] += int(r[ix["Thread Instructions Executed"]] or 0)
        samp[where] += s
        total_inst += n
        total_samp += s
        for c in stall_cols:
            v = int(r[ix[c]] or 0)
            if v:
                stalls[where][c] += v
    print(f"kernel {rows[0][1] if rows and len(rows[0]) > 1 else kid}: {total_inst} warp instr, {total_samp} samples")
    print(f"{'line':28s} {'inst%':>6s} {'thr/inst':>8s} {'samp%':>6s}  top stalls")
    for where in sorted(inst, key=lambda w: -samp[w])[:45]:
        top = sorted(stalls[where].items(), key=lambda kv: -kv[1])[:3]
        tops = " ".join(f"{k[6:]}={v}" for k, v in top)
        print(f"{where:28s} {100 * inst[where] / max(total_inst, 1):6.2f} {thr[where] / max(inst[where], 1):8.1f} "
              f"{100 * samp[where] / max(total_samp, 1):6.2f}  {tops}")


if __name__ == "__main__":
    main()
