"""Top stall locations of one kernel of an ncu report (SASS rows of the source page, grouped by nothing: just the hottest rows).
    python scratch/ncu_hot.py report.ncu-rep kernel_name [rows]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
i_src, i_s = hdr.index("Source"), hdr.index("# Samples")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[2:] if len(r) == len(hdr) and r[i_s].strip().isdigit()]
tot = sum(int(r[i_s] or 0) for r in body)
print("total samples", tot)
agg = {}
for h in stall:
    agg[hdr[h]] = sum(int(r[h] or 0) for r in body)
print({k: round(100 * v / max(1, sum(agg.values())), 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for n, r in sorted(((int(r[i_s] or 0), r) for r in body), key=lambda t: -t[0])[:top]:
    why = sorted(((int(r[h] or 0), hdr[h]) for h in stall), reverse=True)[:2]
    print("%5.1f%%  %-70s %s" % (100 * n / tot, r[i_src][:70], why))
