"""Join an ncu SASS-view source page (ncu -i rep --page source --csv) with nvdisasm -g line info of the same cubin
and aggregate executed warp instructions / stall samples per source line.
usage: ncu_by_line.py source.csv all.sass <kernel substring> [top]"""
import csv, re, sys, collections

src_csv, sass, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
data = [dict(zip(hdr, r)) for r in rows[2:] if len(r) == len(hdr)]
# nvdisasm: instruction sequence of the kernel with the (file, line) in force
lines = open(sass).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('.text.') and kern in l and l.rstrip().endswith(':'))
seq, cur = [], ('?', 0)
for l in lines[start + 1:]:
    if l.startswith('//---') or (l.startswith('.text.') and l.rstrip().endswith(':')):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);', l)
    if m:
        seq.append((int(m.group(1), 16), m.group(2).strip(), cur))
assert len(seq) == len(data), (len(seq), len(data))
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
tot = [0, 0, 0]
for (off, text, loc), d in zip(seq, data):
    n = int(d['Instructions Executed']); s = int(d['# Samples'])
    op = text.split()[1] if text.startswith('@') else text.split()[0]
    f64 = op.startswith(('DFMA', 'DMUL', 'DADD', 'DSETP', 'DMNMX', 'MUFU.RCP64H', 'DSEL'))
    a = agg[loc]; a[0] += n; a[1] += s; a[2] += n if f64 else 0; a[3] += 1
    tot[0] += n; tot[1] += s; tot[2] += n if f64 else 0
src = {}
def line_text(f, ln):
    import os
    for base in ('gym_sbr2_b200/csrc', 'include'):
        p = os.path.join(base, f)
        if os.path.exists(p):
            if p not in src:
                src[p] = open(p).read().splitlines()
            return src[p][ln - 1].strip()[:90] if 0 < ln <= len(src[p]) else ''
    return ''
print('total warp-inst %d, fp64 %d, samples %d, static inst %d' % (tot[0], tot[2], tot[1], len(seq)))
print('%-22s %10s %6s %10s %6s  %s' % ('file:line', 'inst', '%', 'fp64', 'smpl%', 'source'))
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print('%-22s %10d %6.2f %10d %6.2f  %s' % ('%s:%d' % loc, a[0], 100.0 * a[0] / tot[0], a[2], 100.0 * a[1] / max(tot[1], 1), line_text(*loc)))
