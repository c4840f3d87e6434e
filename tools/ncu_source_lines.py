"""hot source lines of one kernel from an .ncu-rep captured with --import-source on (compile with -lineinfo):
   python tools/ncu_source_lines.py gpurun_out/r02_prof_hot.ncu-rep k_l1_build [top]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, fname, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or r[0] == "" or r[0] == "Function Name":
        continue                                  # SASS rows (empty line number) are folded into their source line by ncu
    d = dict(zip(hdr[2:], r[2:]))                 # the first two columns are (Line No, Source)
    try:
        out.append((int(d["# Samples"]), int(d["Instructions Executed"]), float(d["Avg. Threads Executed"] or 0), fname, int(r[0]), r[1].strip()[:105]))
    except (ValueError, KeyError):
        pass
tot, toti = sum(o[0] for o in out) or 1, sum(o[1] for o in out) or 1
print(f"# {kern}: {tot} stall samples, {toti} warp instructions; top {top} source lines by samples")
for o in sorted(out, reverse=True)[:top]:
    print(f"{o[0] * 100 / tot:5.1f}% samples {o[1] * 100 / toti:5.1f}% instr  {o[2]:4.1f} thr/instr  {o[3]}:{o[4]:<4d} {o[5]}")
