"""Per-opcode executed warp-instructions, stall samples and shared-memory wavefronts of one
kernel from `ncu -i rep --page source --csv` (SASS view).  usage: ncu_source_mix.py rep [units]
`units` = number of work units (e.g. tiles x warps) to normalise by."""
import collections
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True,
                     text=True).stdout.splitlines()
rows = list(csv.reader(raw))
hdr = rows[1]
iS, iE, iN = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
iW = hdr.index('L1 Wavefronts Shared')
iG = hdr.index('L1 Tag Requests Global')
stalls = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
cnt, smp, wf, tg = (collections.Counter() for _ in range(4))
st = collections.Counter()
tot = 0
for r in rows[2:]:
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[iS])
    if not m:
        continue
    op = m.group(2)
    n = int(r[iE] or 0)
    cnt[op] += n
    tot += n
    smp[op] += int(r[iN] or 0)
    wf[op] += int(r[iW] or 0)
    tg[op] += int(r[iG] or 0)
    for i in stalls:
        st[hdr[i]] += int(r[i] or 0)
print('total warp-instructions %d = %.1f per unit; samples %d' % (tot, tot / units, sum(smp.values())))
for op, n in cnt.most_common(28):
    print('  %-10s %10d  %7.1f per unit   samples %6d   smem wavefronts %7.1f/unit  global tag req %7.1f/unit'
          % (op, n, n / units, smp[op], wf[op] / units, tg[op] / units))
print('stall samples:', ', '.join('%s %d' % (k, v) for k, v in st.most_common(8)))
