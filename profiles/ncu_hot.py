"""Prints the hottest SASS lines (by warp stall samples) of an ncu report's first kernel.
usage: python profiles/ncu_hot.py gpurun_out/x.ncu-rep [n]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(f"ncu -i {rep} --page source --csv", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1])
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
data = []
for k, r in enumerate(rows[2:]):
    try: data.append((int(r[ci["Warp Stall Sampling (All Samples)"]]), k, r[ci["Source"]].strip(), r[ci["Instructions Executed"]]))
    except Exception: pass
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
for n, k, s, ie in sorted(data, reverse=True)[:top]:
    print(f"{100*n/tot:5.1f}%  #{k:4d} exec={ie:>12s}  {s[:100]}")
