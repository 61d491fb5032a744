"""Per-opcode / per-instruction stall samples from an .ncu-rep source page."""
import collections, csv, subprocess, sys
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (["-k", sys.argv[2]] if len(sys.argv) > 2 else []),
                     capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
data = []
for r in rows[2:]:
    if r and r[0] == 'Kernel Name' and data:
        break                      # several captures in the report: keep the first
    if len(r) == len(hdr) and r[0] != 'Address':
        data.append(r)
ci = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ci['# Samples']]) for r in data)
print(len(data), "SASS instructions;", tot, "samples")
def opc(r):
    s = r[ci['Source']].strip().split()
    op = s[1] if s[0].startswith('@') else s[0]
    return op.split('.')[0]
for reason in ['stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_no_inst', 'stall_mio', 'stall_lg', 'stall_math']:
    t = sum(int(r[ci[reason]]) for r in data)
    by = collections.Counter()
    for r in data:
        by[opc(r)] += int(r[ci[reason]])
    print(f"{reason:16s} {t:7d} ({100*t/tot:4.1f}%)", by.most_common(6))
ex = collections.Counter()
for r in data:
    ex[opc(r)] += int(r[ci['Instructions Executed']])
print("executed:", ex.most_common(18))
for reason in ['stall_long_sb', 'stall_short_sb']:
    print("top", reason)
    for r in sorted(data, key=lambda r: -int(r[ci[reason]]))[:6]:
        print("  ", r[ci[reason]], r[ci['Address']][-5:], r[ci['Source']][:80])
