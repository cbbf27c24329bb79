"""
Per-source-line instruction counts and stall samples of one kernel from an .ncu-rep (works without the GPU):
    python tools/ncu_lines.py REP OBJECT.o KERNEL_MANGLED_SUBSTRING [top]
The SASS page of the report is joined with `nvdisasm -g` line annotations of the same object file by instruction offset.
"""
import csv
import re
import subprocess
import sys
import tempfile
import os
from collections import defaultdict

rep, obj, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
dis = subprocess.run(['nvdisasm', '-g', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
line_of = {}
inside, cur = False, None
for l in dis:
    if l.startswith('.text.'):
        inside = kname in l
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, data = rows[1], rows[2:]
ia, isamp, iaddr = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Address')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
base = int(data[0][iaddr], 16)
per_line = defaultdict(lambda: [0, 0, defaultdict(int)])
tot_i = tot_s = 0
for r in data:
    off = int(r[iaddr], 16) - base
    ln = line_of.get(off, (None, ''))[0]
    a = per_line[ln]
    a[0] += int(r[ia]); a[1] += int(r[isamp])
    for i in stall_cols:
        if r[i] not in ('', '0'):
            a[2][hdr[i]] += int(r[i])
    tot_i += int(r[ia]); tot_s += int(r[isamp])
src = open(sys.argv[5]).read().splitlines() if len(sys.argv) > 5 else None
main_file = os.path.basename(sys.argv[5]) if len(sys.argv) > 5 else None
src_dir = os.path.dirname(sys.argv[5]) if len(sys.argv) > 5 else '.'
_cache = {}
def text_of(key):
    if not key: return ''
    f, n = key
    try:
        if f not in _cache: _cache[f] = open(os.path.join(src_dir, f)).read().splitlines()
        return _cache[f][n - 1].strip()[:90]
    except Exception:
        return ''
print(f'total instructions {tot_i}, samples {tot_s}')
by_file = defaultdict(lambda: [0, 0])
for k, v in per_line.items():
    f = k[0] if k else None
    by_file[f][0] += v[0]; by_file[f][1] += v[1]
for f, v in by_file.items():
    print(f'file {f}: samples {v[1] / tot_s:.3f} instructions {v[0] / tot_i:.3f}')
for ln, a in sorted(per_line.items(), key=lambda kv: -kv[1][1])[:top]:
    st = ', '.join(f'{k[6:]} {v}' for k, v in sorted(a[2].items(), key=lambda kv: -kv[1])[:3])
    text = text_of(ln)
    print(f'{ln}: samples {a[1]:6d} ({a[1] / tot_s:.3f})  inst {a[0] / tot_i:.3f}  [{st}]  {text}')
if len(sys.argv) > 6:      # line ranges "a-b,c-d,..." -> aggregated shares
    for rng in sys.argv[6].split(','):
        a_, b_ = map(int, rng.split('-'))
        s_ = sum(v[1] for k, v in per_line.items() if k is not None and k[0] == main_file and a_ <= k[1] <= b_)
        i_ = sum(v[0] for k, v in per_line.items() if k is not None and k[0] == main_file and a_ <= k[1] <= b_)
        print(f'lines {a_}-{b_}: samples {s_ / tot_s:.3f}  instructions {i_ / tot_i:.3f}  ({src[a_ - 1].strip()[:60] if src else ""})')
