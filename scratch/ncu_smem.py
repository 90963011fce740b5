"""Shared-memory wavefronts per SASS instruction of one kernel (ncu source page): the top consumers.
    python scratch/ncu_smem.py report.ncu-rep kernel_name [rows]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
i_src, i_w, i_id, i_ex = hdr.index("Source"), hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal"), hdr.index("Instructions Executed")
body = [r for r in rows[2:] if len(r) == len(hdr) and r[i_w].strip().isdigit()]
seen, uniq = set(), []
for r in body:
    key = (r[0], r[i_src])
    if key in seen: continue
    seen.add(key); uniq.append(r)
tot = sum(int(r[i_w]) for r in uniq)
print("total shared wavefronts", tot)
agg = {}
for r in uniq:
    op = r[i_src].split()[0] if not r[i_src].startswith("@") else r[i_src].split()[1]
    a = agg.setdefault(op, [0, 0, 0]); a[0] += int(r[i_w]); a[1] += int(r[i_id] or 0); a[2] += int(r[i_ex] or 0)
for op, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:12]:
    print("%-28s wavefronts %12d (%.1f%%) ideal %12d  executed %d" % (op, a[0], 100 * a[0] / max(tot, 1), a[1], a[2]))
for r in sorted(uniq, key=lambda r: -int(r[i_w]))[:top]:
    print("%5.1f%%  %-64s wf %s ideal %s exec %s" % (100 * int(r[i_w]) / max(tot, 1), r[i_src][:64], r[i_w], r[i_id], r[i_ex]))
