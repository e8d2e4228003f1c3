"""Per CUDA source line: share of the warp-stall samples of one kernel, from
    ncu -i REPORT --page source --csv --print-source sass,cuda --kernel-name NAME > lines.csv
    python tools/ncu_line_summary.py lines.csv [top_n]"""
import collections
import csv
import sys


def main(path, top=30):
    rows = list(csv.reader(open(path)))
    out = []
    cur_file = ""
    hdr = None
    per = collections.OrderedDict()
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            i_s = [i for i, c in enumerate(hdr) if c == "# Samples"][0]
            i_x = [i for i, c in enumerate(hdr) if c == "Instructions Executed"][0]
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        try:
            s = float(r[i_s] or 0); x = float(r[i_x] or 0)
        except ValueError:
            continue
        key = (cur_file, r[0])
        if key not in per:
            per[key] = [0.0, 0.0, r[1]]
        per[key][0] += s; per[key][1] += x
    S = sum(v[0] for v in per.values()) or 1.0
    X = sum(v[1] for v in per.values()) or 1.0
    for (f, ln), (s, x, src) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * s / S:5.1f}% samples {100 * x / X:5.1f}% instr  {f}:{ln:>4}  {src.strip()[:110]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
