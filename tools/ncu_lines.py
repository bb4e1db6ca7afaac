"""Aggregate an `ncu --page source --csv` (SASS view) export by CUDA source line.
usage: ncu_lines.py <src.csv> <object.o> <kernel-substring> [top]
The SASS -> line map comes from `nvdisasm -g` on the cubin inside the object file."""
import csv
import re
import subprocess
import sys
import tempfile
from collections import defaultdict
from pathlib import Path

src_csv, obj, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", str(Path(obj).resolve())], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = next(Path(tmp).glob("*.cubin"))
dis = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True, text=True).stdout.splitlines()
addr2line = {}
inside, cur = False, None
for ln in dis:
    if ln.startswith("//---") and ".text." in ln:
        inside = kern in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        # keep the innermost user-code line of an inlined chain
        cur = (Path(m.group(1)).name, int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        addr2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(src_csv)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ci = {n: i for i, n in enumerate(hdr)}
base = None
per_line = defaultdict(lambda: [0, 0, defaultdict(int)])
total = 0
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
for r in rows[h + 1:]:
    if len(r) < len(hdr):
        continue
    a = int(r[ci["Address"]], 16)
    if base is None:
        base = a
    off = a - base
    smp = int(r[ci["# Samples"]] or 0)
    ex = int(r[ci["Instructions Executed"]] or 0)
    total += smp
    line, sass = addr2line.get(off, (None, r[ci["Source"]]))
    e = per_line[line]
    e[0] += smp
    e[1] += ex
    for n in stall_cols:
        v = int(r[ci[n]] or 0)
        if v:
            e[2][n] += v
print(f"total samples {total}")
for line, (smp, ex, st) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
    tops = ", ".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:4])
    print(f"{100.0 * smp / max(total, 1):5.1f}%  {smp:7d} smp  {ex:10d} inst  {line}  [{tops}]")
