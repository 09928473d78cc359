"""Per source line of one kernel from `ncu --page source --print-source cuda,sass --csv`: stall samples WITHOUT the barrier
waits (so the idle warps of a phase do not drown the working ones), warp instructions executed, average active threads,
and the main stall reasons.  usage: ncu_lines2.py report.ncu-rep kernel_regex [file_substring] [lo-hi]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
fsub = sys.argv[3] if len(sys.argv) > 3 else ""
lo, hi = (map(int, sys.argv[4].split("-")) if len(sys.argv) > 4 else (0, 10**9))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = ""; hdr = None; agg = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 4 and r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None or len(r) < 40 or not r[0]: continue
    try: ln = int(r[0])
    except ValueError: continue
    if fsub not in cur or not (lo <= ln <= hi): continue
    def g(k):
        try: return float(r[hdr[k]])
        except (ValueError, KeyError): return 0.0
    a = agg.setdefault((cur, ln), dict(src=r[1].strip()[:90], s=0, b=0, inst=0, thr=0, lsb=0, ssb=0, mio=0, lg=0, wait=0, br=0, ns=0))
    a["s"] += g("# Samples"); a["b"] += g("stall_barrier"); a["inst"] += g("Instructions Executed"); a["thr"] += g("Thread Instructions Executed")
    a["lsb"] += g("stall_long_sb"); a["ssb"] += g("stall_short_sb"); a["mio"] += g("stall_mio"); a["lg"] += g("stall_lg"); a["wait"] += g("stall_wait")
    a["br"] += g("stall_branch_resolving"); a["ns"] += g("stall_not_selected") + g("stall_selected")
tot = sum(a["s"] - a["b"] for a in agg.values())
print("non-barrier samples", tot, " barrier samples", sum(a["b"] for a in agg.values()))
print("%5s %7s %6s %9s %5s | %6s %6s %6s %6s %6s %6s" % ("line", "work", "%", "warpinst", "thr", "longsb", "shrtsb", "mio", "lg", "wait", "sel"))
for (f, ln), a in sorted(agg.items(), key=lambda kv: kv[0]):
    w = a["s"] - a["b"]
    if w < tot * 0.002 and a["inst"] < 1000: continue
    print("%5d %7d %5.1f%% %9d %5.1f | %6d %6d %6d %6d %6d %6d  %s" % (ln, w, 100.0 * w / max(tot, 1), a["inst"], a["thr"] / max(a["inst"], 1),
          a["lsb"], a["ssb"], a["mio"], a["lg"], a["wait"], a["ns"], a["src"]))
