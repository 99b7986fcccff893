"""Scratch: basic-block execution profile from an ncu report (SASS page). usage: ncu_blocks.py report.ncu-rep [min_Mwarpinst]"""
import csv, re, subprocess, sys, io
rep = sys.argv[1]; thr = float(sys.argv[2]) * 1e6 if len(sys.argv) > 2 else 15e6
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
his = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hi = his[which]; end = his[which + 1] - 1 if which + 1 < len(his) else len(rows)
hdr = rows[hi]; data = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
ci = {n: i for i, n in enumerate(hdr)}
base = int(data[0][0], 16)
ins = [dict(off=int(r[0], 16) - base, ie=int(r[ci['Instructions Executed']]), te=int(r[ci['Thread Instructions Executed']]),
            smp=int(r[ci['# Samples']]), src=r[1].strip()) for r in data]
tot = sum(x['ie'] for x in ins); tt = sum(x['te'] for x in ins); ts = sum(x['smp'] for x in ins)
print('total warp inst %.4g  thread inst %.4g  avg threads %.2f  samples %d' % (tot, tt, tt / tot, ts))
blocks = []; cur = []
for x in ins:
    if cur and (x['ie'] != cur[-1]['ie'] or re.search(r'\b(BSYNC|BSSY)\b', x['src'])):
        blocks.append(cur); cur = []
    cur.append(x)
    if re.search(r'\b(BRA|CALL|EXIT|RET)\b', x['src']): blocks.append(cur); cur = []
if cur: blocks.append(cur)
for b in blocks:
    w = sum(x['ie'] for x in b); t = sum(x['te'] for x in b); s = sum(x['smp'] for x in b)
    if w < thr: continue
    ops = {}
    for x in b:
        m = x['src'].split(); k = (m[1] if m[0].startswith('@') else m[0]).split('.')[0]
        ops[k] = ops.get(k, 0) + 1
    top = ' '.join('%s:%d' % kv for kv in sorted(ops.items(), key=lambda kv: -kv[1])[:7])
    print('%05x-%05x n=%3d exec=%7.2fM warpinst=%7.1fM (%4.1f%%) thr=%4.1f smp=%5.1f%% | %s' %
          (b[0]['off'], b[-1]['off'], len(b), b[0]['ie'] / 1e6, w / 1e6, 100 * w / tot, t / max(w, 1), 100.0 * s / ts, top))
