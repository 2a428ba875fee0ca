"""Aggregate an `ncu --page source --csv` export: stall samples per SASS opcode and the
hottest instructions; shared-memory wavefronts per LDS/STS instruction."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
def num(r, k):
    try: return float(r[col[k]])
    except Exception: return 0.0
tot = sum(num(r, '# Samples') for r in data)
by = collections.Counter(); inst = collections.Counter()
stall_keys = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
stalls = collections.Counter()
for r in data:
    op = r[col['Source']].split()[0] if r[col['Source']].split() else ''
    if op.startswith('@'): op = r[col['Source']].split()[1]
    by[op] += num(r, '# Samples'); inst[op] += num(r, 'Instructions Executed')
    for k in stall_keys: stalls[k] += num(r, k)
print("total samples", tot)
print("stall totals:", ", ".join(f"{k[6:]}={int(v)}" for k, v in stalls.most_common(12)))
print("%-34s %10s %7s %14s" % ("opcode", "samples", "%", "warp-instrs"))
for op, s in by.most_common(22):
    print("%-34s %10d %6.1f%% %14d" % (op, s, 100 * s / tot, inst[op]))
print("\nhottest instructions:")
for r in sorted(data, key=lambda r: -num(r, '# Samples'))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    top = sorted(((num(r, k), k[6:]) for k in stall_keys), reverse=True)[:3]
    print("%6d  %-64s %s" % (num(r, '# Samples'), r[col['Source']].strip()[:64], " ".join(f"{k}:{int(v)}" for v, k in top if v)))
print("\nshared-memory instructions (wavefronts / ideal per warp-instr):")
seen = 0
for r in data:
    w = num(r, 'L1 Wavefronts Shared')
    if w and seen < 40:
        n = num(r, 'Instructions Executed')
        print("  %-56s wf/instr %.2f ideal %.2f" % (r[col['Source']].strip()[:56], w / max(n, 1), num(r, 'L1 Wavefronts Shared Ideal') / max(n, 1)))
        seen += 1
