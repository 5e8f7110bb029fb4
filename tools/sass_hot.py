"""Static look at the hot loop of a kernel in libfds_b200.so (development aid, no GPU needed).

    python tools/sass_hot.py <mangled-name-substring> [--list]

Finds the loop with the most FP64 instructions (the unchecked two-block body of the step kernel), walks it
once skipping every span a conditional FORWARD branch jumps over (rare paths: node update, copy issue, blocking
waits) and prints the instruction mix of what is left: 2 issue cycles per FP64 instruction + 1 per other
instruction is the kernel's time (profiles/r2a_step_ncu.txt: 2 * FP64 + others = measured cycles).
"""
import collections
import re
import subprocess
import sys

import os
so = os.environ.get('FDS_SASS_BIN', 'fds_b200/libfds_b200.so')
pat = sys.argv[1]
names = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
funcs = re.findall(r'Function : (\S+)', names)
cands = [f for f in funcs if pat in f]
assert cands, 'no function matches'
fn = cands[0]
txt = subprocess.run(['cuobjdump', '-sass', '-fun', fn, so], capture_output=True, text=True).stdout
ins = []
for l in txt.splitlines():
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(ins)}


def opc(t):
    t = re.sub(r'^@!?U?P\d+\s+', '', t)
    return t.split()[0].split('.')[0]


FP = ('DADD', 'DMUL', 'DFMA')
MINFP = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 600
loops = []
for i, (a, t) in enumerate(ins):
    if 'BRA' in t:
        m = re.search(r'0x([0-9a-f]+)', t)
        if m and int(m.group(1), 16) <= a and int(m.group(1), 16) in addr:
            loops.append((addr[int(m.group(1), 16)], i))
best = None
for lo, hi in loops:
    nfp = sum(1 for _, t in ins[lo:hi + 1] if opc(t) in FP)
    size = hi - lo
    if nfp >= MINFP and (best is None or size < best[2]):
        best = (lo, hi, size, nfp)
lo, hi = best[0], best[1]
hot, i = [], lo
while i <= hi:
    a, t = ins[i]
    hot.append((a, t))
    if 'BRA' in t and t.startswith('@'):
        m = re.search(r'0x([0-9a-f]+)', t)
        tgt = int(m.group(1), 16) if m else None
        if tgt in addr and a < tgt <= ins[hi][0]:
            i = addr[tgt]
            continue
    i += 1
mix = collections.Counter(opc(t) for _, t in hot)
nfp = sum(mix[o] for o in FP)
oth = len(hot) - nfp
print('%s\nloop %#x..%#x: %d instructions static, hot path %d = %d FP64 + %d others -> %d issue cycles'
      % (fn, ins[lo][0], ins[hi][0], hi - lo + 1, len(hot), nfp, oth, 2 * nfp + oth))
print(sorted(((k, v) for k, v in mix.items() if k not in FP), key=lambda kv: -kv[1]))
if '--list' in sys.argv:
    for a, t in hot:
        if opc(t) not in FP:
            print('%05x  %s' % (a, t))
