"""Per-kernel SASS evidence from the built library (no GPU needed):
    python tools/sass_summary.py > profiles/r02_sass_summary.txt
counts the mnemonics that prove which hardware paths a kernel uses -- tcgen05 MMA (UTCHMMA), TMEM loads (LDTM),
bulk TMA copies (UBLKCP), mbarrier operations (SYNCS), cp.async (LDGSTS), 128-bit global / shared accesses, atomics."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "audio-analyzer-omega_b200", "omega4_b200", "libomega4_cuda.so")
WANT = ["UTCHMMA", "UTCIMMA", "UTCBAR", "LDTM", "UBLKCP", "SYNCS", "LDGSTS", "LDG.E.128", "LDG.E.64", "LDG.E", "STG.E.128", "STG.E.64",
        "STG.E", "LDS.128", "LDS.64", "LDS", "STS.128", "STS.64", "STS", "LD.E", "ST.E", "REDG", "ATOMG", "REDUX", "SHFL", "BAR.SYNC",
        "FFMA", "FADD", "FMUL", "DFMA", "MUFU", "F2F"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}   (architectures in the fat binary: {', '.join(arch)})")
    cur, per = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
        if m:
            per[cur]["_total"] += 1
            op = m.group(1)
            for w in WANT:
                if op == w or op.startswith(w + "."):
                    per[cur][w] += 1
                    break
    for name, c in per.items():
        short = re.sub(r"^void ", "", name).replace("o4::", "")
        print(f"\n{short}   ({c['_total']} SASS instructions)")
        print("   " + "  ".join(f"{w} {c[w]}" for w in WANT if c[w]))


if __name__ == "__main__":
    main()
