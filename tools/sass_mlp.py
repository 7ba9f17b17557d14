"""Static look at gather batching in a kernel's SASS: for every 128-bit global load, how many
further global loads are issued before the first instruction that reads its destination.
usage: python tools/sass_mlp.py object.o kernel-name-substring"""
import re, subprocess, sys
obj, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
ins, on = [], False
for line in txt.splitlines():
    if 'Function :' in line:
        on = pat in line
        continue
    if on:
        m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(.*?);', line)
        if m:
            ins.append(m.group(1).strip())
loads = [i for i, s in enumerate(ins) if re.search(r'\bLDG\.E\.(EF\.)?128', s)]
res = []
for i in loads:
    m = re.search(r'LDG\S*\s+R(\d+)', ins[i])
    d = int(m.group(1))
    regs = {'R%d' % (d + k) for k in range(4)}
    for j in range(i + 1, min(i + 3000, len(ins))):
        body = ins[j].split(None, 1)[1] if ' ' in ins[j] else ''
        # source operands = everything after the first comma (rough) or any operand for stores
        toks = set(re.findall(r'R\d+', body))
        srcs = set(re.findall(r'R\d+', body.split(',', 1)[1])) if ',' in body and not ins[j].startswith(('STG', 'STS', 'STL')) else toks
        if regs & srcs:
            between = sum(1 for k in loads if i < k < j)
            res.append((i, j - i, between))
            break
print('%d instructions, %d wide global loads' % (len(ins), len(loads)))
import collections
hist = collections.Counter(min(b, 15) for _, _, b in res)
print('loads issued after a load and before its first use (15 = 15+):')
print('  ' + '  '.join('%d:%d' % (k, hist[k]) for k in sorted(hist)))
print('  mean %.1f' % (sum(b for _, _, b in res) / max(1, len(res))))
