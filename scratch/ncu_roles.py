import csv, subprocess, sys, re
path = sys.argv[1]
src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines())); hdr = rows[1]
iS = hdr.index('Source'); iN = hdr.index('# Samples'); iE = hdr.index('Instructions Executed')
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
ins = rows[2:]
# split into regions at role markers
marks = {'UTMALDG': 'TMA', 'UTCHMMA': 'MMA', 'STTM': 'Aconv', 'LDTM': 'epi', 'UTMASTG': 'epi'}
role_at = [None] * len(ins)
for i, r in enumerate(ins):
    for k, v in marks.items():
        if k in r[iS]: role_at[i] = v
# print a compact listing: every instruction with >=0.3% samples, with nearest marker role before/after
tot = sum(int(r[iN]) for r in ins if r[iN].isdigit())
print('total', tot)
last = None
for i, r in enumerate(ins):
    if role_at[i]: last = role_at[i]
    n = int(r[iN]) if r[iN].isdigit() else 0
    if n >= tot * 0.004:
        nxt = next((role_at[k] for k in range(i, len(ins)) if role_at[k]), None)
        st = {hdr[k][6:]: int(r[k]) for k in stall if r[k] not in ('', '0') and int(r[k]) > n * 0.15}
        print(f"{i:5d} {n:6d} {100*n/tot:5.1f}% exec {r[iE]:>9s} prev={last} next={nxt}  {r[iS][:60]:60s} {st}")
