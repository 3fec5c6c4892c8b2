#!/usr/bin/env python
"""Where the executed instructions of one short_kernel launch go, by PHASE of the kernel.

    python tools/ncu_regions.py REPORT.ncu-rep DEMANGLED_SUBSTRING MANGLED_SUBSTRING [host.o]

e.g. "(int)15, (int)256" "short_kernelIjLi15ELi256ELi2ELb0ELi32ELb1ELb1ELb0ELb0E" for the fused ASCII launch.

Joins `ncu --page source --csv` (SASS rows: executed warp / thread instructions, stall samples) with
the inline chains of `nvdisasm -gi` (built with -lineinfo) by instruction offset.  Every instruction is
attributed to the OUTERMOST frame of its chain (a line of the kernel body in short_kernel.cuh) and
through that to the phase whose marker comment ("---------------- N. name") precedes the line; inside
the compute phase the totals are also split by the called row function's own lines.
"""
import csv
import io
import re
import subprocess
import sys
import tempfile
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
KERNEL_SRC = ROOT / "polars-strsim_b200" / "csrc" / "short_kernel.cuh"


def chains(obj: Path, needle: str):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", str(obj.resolve())], cwd=tmp, check=True, capture_output=True)
    cubin = next(Path(tmp).glob("*.cubin"))
    txt = subprocess.run(["nvdisasm", "-gi", "-c", str(cubin)], capture_output=True, text=True).stdout
    out, cur, active, fresh = {}, [], False, True
    for line in txt.splitlines():
        if line.startswith("//---") and ".text." in line:
            active = needle in line
            cur = []
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
        if m:
            if fresh:
                cur, fresh = [], False
            cur.append((Path(m.group(1)).name, int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            out[int(m.group(1), 16)] = (list(cur), m.group(2))
            fresh = True
    return out


def phases():
    """(first line, name) of every phase marker in the kernel body, in order."""
    marks = [(401, "0. prologue / tile loop")]
    for no, line in enumerate(KERNEL_SRC.read_text().splitlines(), 1):
        m = re.match(r"\s*// -{10,} (\d\. [a-z]+)", line)
        if m and no > 401:
            marks.append((no, m.group(1)))
    return marks


def main():
    rep, kid, needle = sys.argv[1], sys.argv[2], sys.argv[3]
    obj = Path(sys.argv[4] if len(sys.argv) > 4 else ROOT / "polars-strsim_b200/csrc/host.o")
    ch = chains(obj, needle)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    pick = next(i for i in starts if kid in rows[i][1])
    end = next((j for j in starts if j > pick), len(rows))
    rows = rows[pick:end]
    print(rows[0][1])
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    ix = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[hdr_i + 1:] if r and re.match(r"^(0x)?[0-9a-f]+$", r[0])]
    base = int(body[0][0], 16)
    marks = phases()

    def phase_of(line):
        name = marks[0][1]
        for first, nm in marks:
            if line >= first:
                name = nm
        return name

    by_phase = defaultdict(lambda: [0, 0, 0])
    by_line = defaultdict(lambda: [0, 0, 0])
    by_inner = defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for r in body:
        off = int(r[0], 16) - base
        chain, _ = ch.get(off, ([], ""))
        n = int(r[ix["Instructions Executed"]] or 0)
        t = int(r[ix["Thread Instructions Executed"]] or 0)
        s = int(r[ix["# Samples"]] or 0)
        outer = next((c for c in reversed(chain) if c[0] == "short_kernel.cuh" and c[1] >= 401), None)
        ph = phase_of(outer[1]) if outer else "?"
        inner = chain[0] if chain else ("?", 0)
        # the frame just inside the kernel body: which callee this instruction belongs to
        callee = "-"
        if outer and chain.index(outer) > 0:
            c = chain[chain.index(outer) - 1]
            callee = f"{c[0]}:{c[1]}"
        for d, k in ((by_phase, ph), (by_line, (ph, outer[1] if outer else 0, callee)),
                     (by_inner, (ph, f"{inner[0]}:{inner[1]}"))):
            d[k][0] += n
            d[k][1] += t
            d[k][2] += s
        tot[0] += n
        tot[1] += t
        tot[2] += s
    print(f"total: {tot[0]} warp instr, {tot[1]} thread instr ({tot[1] / max(tot[0], 1):.1f} lanes), {tot[2]} samples")
    print("\n== by phase ==")
    for ph, (n, t, s) in sorted(by_phase.items()):
        print(f"{ph:28s} inst {100 * n / tot[0]:6.2f}%  lanes {t / max(n, 1):5.1f}  samples {100 * s / max(tot[2], 1):6.2f}%")
    print("\n== by kernel-body line / callee (>= 0.4 % of instructions) ==")
    for (ph, line, callee), (n, t, s) in sorted(by_line.items(), key=lambda kv: (kv[0][0], kv[0][1], kv[0][2])):
        if n >= 0.004 * tot[0] or s >= 0.01 * tot[2]:
            print(f"{ph:16s} :{line:<4d} {callee:26s} inst {100 * n / tot[0]:6.2f}%  lanes {t / max(n, 1):5.1f}  "
                  f"samples {100 * s / max(tot[2], 1):6.2f}%")
    print("\n== by innermost line (>= 0.8 % of instructions) ==")
    for (ph, where), (n, t, s) in sorted(by_inner.items(), key=lambda kv: -kv[1][0]):
        if n >= float(__import__("os").environ.get("MIN_INNER", "0.008")) * tot[0]:
            print(f"{ph:16s} {where:28s} inst {100 * n / tot[0]:6.2f}%  lanes {t / max(n, 1):5.1f}  "
                  f"samples {100 * s / max(tot[2], 1):6.2f}%")


if __name__ == "__main__":
    main()
