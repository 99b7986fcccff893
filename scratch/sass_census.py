"""Scratch: instruction census per kernel from cuobjdump -sass (profiles/r01_sass_census.txt)."""
import re, subprocess, sys, collections
lib = sys.argv[1] if len(sys.argv) > 1 else 'chroma_lite_b200/libchroma_b200.so'
txt = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
keys = ['UBLKCP', 'SYNCS', 'LDG.E.128', 'LDG.E.64', 'LDG.E', 'LDS', 'STS', 'LDL', 'STL', 'PRMT', 'VOTE', 'SHFL', 'CREDUX', 'REDUX', 'ATOMG', 'RED', 'DFMA',
        'MUFU.RCP64H', 'MUFU', 'FFMA', 'IMAD', 'BAR', 'CALL', 'HMMA', 'UTCMMA']
print('# cuobjdump -sass %s (sm_100a), instruction census per kernel' % lib)
print('# UBLKCP = cp.async.bulk (TMA 1-D bulk copy of the optical tables into shared memory), SYNCS = mbarrier,')
print('# LDG.E.128(.CONSTANT) = 128-bit read-only node / triangle fetches, PRMT = near/far plane select, VOTE/SHFL/CREDUX = warp')
print('# aggregation (queue append, refill, ray splitting, warp-cooperative reduction), ATOMG/RED = queue cursors and DAQ atomics,')
print('# DFMA/MUFU.RCP64H = the double-precision reciprocal of the reference triangle test, HMMA/UTC*MMA: none (no dense contraction).')
cur, counts = None, None
def flush():
    if cur and not cur.startswith('_ZN3cub'):
        tot = sum(counts.values())
        parts = []
        for k in keys:
            n = sum(v for m, v in counts.items() if m.startswith(k) and not any(m.startswith(k2) for k2 in keys if len(k2) > len(k) and k2.startswith(k)))
            if n: parts.append('%s=%d' % (k, n))
        print(cur); print('    instructions=%d  %s' % (tot, '  '.join(parts)))
for line in txt.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        flush(); cur = m.group(1); counts = collections.Counter(); continue
    m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)', line)
    if m and cur: counts[m.group(1)] += 1
flush()
