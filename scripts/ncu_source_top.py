"""Top source lines of an ncu report by stall samples / executed instructions:
   python scripts/ncu_source_top.py gpurun_out/x.ncu-rep [n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
cur, hdr, out = None, None, []
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"):
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        smp = int(r[hdr.index("# Samples")] or 0)
        ins = int(r[hdr.index("Instructions Executed")] or 0)
        out.append((smp, ins, cur, int(r[0]), r[1].strip()[:110]))
ts, ti = sum(o[0] for o in out), sum(o[1] for o in out)
print("total samples %d, warp instructions %d" % (ts, ti))
byfile = {}
for o in out:
    a = byfile.setdefault(o[2], [0, 0])
    a[0] += o[0]
    a[1] += o[1]
for f, v in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print("  %-22s samples %5.1f %%  instructions %5.1f %%" % (f, 100.0 * v[0] / max(ts, 1), 100.0 * v[1] / max(ti, 1)))
for o in sorted(out, key=lambda o: -o[0])[:n]:
    print("%5.1f%% smp %5.1f%% ins  %s:%d  %s" % (100.0 * o[0] / max(ts, 1), 100.0 * o[1] / max(ti, 1), o[2], o[3], o[4]))
