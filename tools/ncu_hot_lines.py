"""Top source lines by warp-stall samples from `ncu --page source --print-source cuda,sass --csv`."""
import csv
import subprocess
import sys


def main(rep, kernel_regex, top=40):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                          "--kernel-name", "regex:" + kernel_regex], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file = ""
    agg = {}
    total = 0
    hdr = None
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if len(r) > 4 and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < 6:
            continue
        if r[0]:      # a source line row: aggregated samples
            try:
                n = int(r[4])
            except ValueError:
                continue
            key = (cur_file, int(r[0]), r[1].strip()[:110])
            agg[key] = agg.get(key, 0) + n
            total += n
    print("total samples", total)
    for (f, ln, src), n in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
        print("%6d %5.1f%%  %s:%d  %s" % (n, 100.0 * n / max(total, 1), f, ln, src))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
