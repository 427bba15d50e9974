"""Per-kernel SASS mnemonic census of the shipped library (cuobjdump -sass): which kernels use tcgen05 / TMEM / TMA.
    python tools/sass_evidence.py > profiles/r02_sass_evidence.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
LIB = os.path.join(ROOT, "simple_tip_b200", "libb200tip.so")
COLS = ["UTCHMMA", "UTCCP", "LDTM", "UTCBAR", "UTMALDG", "SYNCS", "REDUX", "MUFU.EX2", "RED/ATOM", "LD/ST .SYS"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = {}
    cur = None
    count = collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            count[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if cur and m:
            op = m.group(1)
            c = count[cur]
            c["n"] += 1
            for key in ("UTCHMMA", "UTCCP", "LDTM", "UTCBAR", "UTMALDG", "SYNCS", "REDUX", "MUFU.EX2"):
                if op.startswith(key):
                    c[key] += 1
            if op.startswith("RED") and not op.startswith("REDUX") or op.startswith("ATOM"):
                c["RED/ATOM"] += 1
            if ".SYS" in op and (op.startswith("LD") or op.startswith("ST")):
                c["LD/ST .SYS"] += 1
    dem = subprocess.run(["c++filt"] + list(count), capture_output=True, text=True).stdout.splitlines()
    for k, d in zip(count, dem):
        names[k] = re.sub(r"\(.*", "", d).replace("void ", "")
    print("# SASS evidence (`cuobjdump -sass simple_tip_b200/libb200tip.so`, sm_100a), round 2\n")
    print("Per kernel: instruction count and the mnemonics that show tcgen05 / TMEM / TMA use (B200_PROFILING.md: "
          "`tcgen05.mma` -> UTCHMMA, `tcgen05.cp` -> UTCCP, `tcgen05.ld` -> LDTM, `tcgen05.commit` -> UTCBAR, TMA -> UTMALDG, mbarrier -> SYNCS); "
          "`LD/ST .SYS` = system-scope loads / stores (the peer-memory exchange of csrc/shard.cu).  "
          "Regenerate: `python tools/sass_evidence.py`.\n")
    print("| kernel | SASS instr | " + " | ".join(COLS) + " |")
    print("|---|---:|" + "---:|" * len(COLS))
    tot = collections.Counter()
    for k, c in count.items():
        print(f"| `{names[k]}` | {c['n']} | " + " | ".join(str(c[x]) for x in COLS) + " |")
        tot.update(c)
    print(f"| **total ({len(count)} kernels)** | {tot['n']} | " + " | ".join(str(tot[x]) for x in COLS) + " |")


if __name__ == "__main__":
    sys.exit(main())
