"""Attribute ncu per-SASS-instruction counters to CUDA source lines.
usage: sass_lines.py <report.ncu-rep> <kernel mangled-name substring> <nvdisasm -g -c listing> <source.cu>"""
import collections
import csv
import re
import subprocess
import sys

rep, kname, listing, src = sys.argv[1:5]
# offset -> line from nvdisasm
lines = open(listing).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l and l.rstrip().endswith(":"))
off2line, cur = {}, None
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.startswith("//-----"):
        if off2line:
            break
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m:
        cur = (m.group(1).rsplit("/", 1)[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m:
        off2line[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Address")
ia = hdr.index("Instructions Executed")
isamp = hdr.index("# Samples") if "# Samples" in hdr else hdr.index("Warp Stall Sampling (All Samples)")
want = sys.argv[5] if len(sys.argv) > 5 else ""
body, take = [], False
for r in rows:
    if r and r[0] == "Kernel Name":
        take = (want in r[1]) and not body
        continue
    if take and len(r) > 6 and r[0].startswith("0x"):
        body.append(r)
base = int(body[0][0], 16)
by_exec, by_samp = collections.Counter(), collections.Counter()
for r in body:
    ln = off2line.get(int(r[0], 16) - base)
    by_exec[ln] += int(r[ia])
    by_samp[ln] += int(r[isamp])
import os
_src_cache = {}
def text(key):
    if not key:
        return "?"
    f, ln = key
    if f not in _src_cache:
        path = os.path.join(os.path.dirname(src), f)
        _src_cache[f] = open(path).read().splitlines() if os.path.exists(path) else []
    lines_ = _src_cache[f]
    return f"{f}:{ln}: " + (lines_[ln - 1].strip()[:100] if 0 < ln <= len(lines_) else "")
te, ts = sum(by_exec.values()), (sum(by_samp.values()) or 1)
print(f"total executed {te}  samples {ts}")
print("--- by executed instructions")
for ln, c in by_exec.most_common(28):
    print(f"{100*c/te:5.1f}% exec {100*by_samp[ln]/ts:5.1f}% samp  {text(ln)}")
print("--- by stall samples")
for ln, c in by_samp.most_common(16):
    print(f"{100*c/ts:5.1f}% samp {100*by_exec[ln]/te:5.1f}% exec  {text(ln)}")
