"""Summarise an ncu --set full report of gconv_stack_tc_kernel: wait sites and top stalls (development aid)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ia, isrc, isamp, iexec = hdr.index('Address'), hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
acc = {}
order = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    try: s = int(r[isamp]); e = int(r[iexec])
    except ValueError: continue
    a = r[ia]
    if a not in acc:
        acc[a] = [s, e, r[isrc], r]; order.append(a)
tot = sum(v[0] for v in acc.values())
print("total samples", tot, "instructions", len(acc))
top = sorted(acc.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]
for a, (s, e, src, r) in top:
    st = sorted([(int(r[i]) if r[i] else 0, hdr[i]) for i in stall_cols], reverse=True)[:2]
    print(f"{s:7d} {100*s/tot:5.1f}% exec={e:9d} {a[-5:]} {src[:64]:64s} {st}")
