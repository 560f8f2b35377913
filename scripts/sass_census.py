"""SASS opcode census of libjoxsz_b200.so per kernel: which Blackwell / tensor / FP64 / copy instructions each kernel
really contains (cuobjdump -sass).  Writes profiles/<tag>_sass_census.txt.   python scripts/sass_census.py r02"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "joxsz_b200", "libjoxsz_b200.so")
WATCH = ["UTCIMMA", "UTCHMMA", "UTCQMMA", "UTCMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS",
         "DMMA", "IMMA", "HMMA", "DFMA", "DADD", "DMUL", "MUFU", "LDGSTS", "LDS", "STS", "LDG", "STG", "BAR", "SHFL"]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for ln in sass.splitlines():
        m = re.match(r"\s+Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m and cur:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    kernels[cur][w] += 1
                    break
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), stdout=subprocess.PIPE, text=True).stdout.splitlines()
    out = [f"SASS opcode census of joxsz_b200/libjoxsz_b200.so (sm_100a), `cuobjdump -sass`, counts of static instructions",
           "columns: " + " ".join(WATCH), ""]
    for (name, c), dm in zip(kernels.items(), demangle):
        short = re.sub(r"\(anonymous namespace\)::", "", dm)
        short = re.sub(r"\(.*", "", short)
        out.append(f"{short}   [{c['_total']} instructions]")
        out.append("    " + "  ".join(f"{w}={c[w]}" for w in WATCH if c[w]))
    path = os.path.join(ROOT, "profiles", f"{tag}_sass_census.txt")
    open(path, "w").write("\n".join(out) + "\n")
    print(path)


if __name__ == "__main__":
    main()
