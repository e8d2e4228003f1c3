"""Summarise `ncu --page source --csv` output: stall shares and opcode mix of one kernel."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    idx = {k: i for i, k in enumerate(hdr)}
    stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    tot, ops, samp = collections.Counter(), collections.Counter(), collections.Counter()
    n = 0
    for r in rows[h + 1:]:
        if len(r) < len(hdr) or r[0] == "Address":
            continue
        n += 1
        for s in stalls:
            try:
                tot[s] += int(r[idx[s]])
            except ValueError:
                pass
        toks = r[idx["Source"]].split()
        op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
        op = op.split(".")[0]
        try:
            ops[op] += int(r[idx["Instructions Executed"]])
            samp[op] += int(r[idx["# Samples"]])
        except ValueError:
            pass
    print(n, "SASS instructions")
    S = sum(tot.values()) or 1
    for k, v in tot.most_common(9):
        print(f"  {k:25s} {v:9d} {100 * v / S:5.1f}%")
    T = sum(ops.values()) or 1
    for k, v in ops.most_common(16):
        print(f"  {k:10s} exec {v:13d} {100 * v / T:5.1f}%   samples {samp[k]}")


if __name__ == "__main__":
    main(sys.argv[1])
