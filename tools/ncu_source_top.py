"""Top stall locations (SASS) of one kernel in an ncu report: python tools/ncu_source_top.py rep kernel_regex [launch_idx]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
idx = sys.argv[3] if len(sys.argv) > 3 else "1"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name-base", "demangled", "-k", "regex:" + rx,
                      "--launch-skip", str(int(idx) - 1), "--launch-count", "1"], capture_output=True, text=True).stdout.splitlines()
rows = [r for r in csv.reader(out[1:]) if len(r) > 10]
hdr = rows[0]
rows = [hdr] + [r for r in rows[1:] if len(r) == len(hdr) and r[0] != "Address"]
ci = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ci["# Samples"]] or 0) for r in rows[1:])
print("total samples", tot)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ci[h]] or 0) for r in rows[1:]) for h in stalls}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
top = sorted(rows[1:], key=lambda r: -int(r[ci["# Samples"]] or 0))[:28]
for r in top:
    st = sorted(((h, int(r[ci[h]] or 0)) for h in stalls), key=lambda kv: -kv[1])[:2]
    print("%6s  %-70s %s" % (r[ci["# Samples"]], r[ci["Source"]].strip()[:70], st))
