"""Evidence helper (not a test): per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths of
libmpvae_b200.so (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG / UBLKCP, ...).

    python tests/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mpvae-1_b200", "libmpvae_b200.so")
WATCH = ["UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "UTCBAR", "USETMAXREG", "SYNCS", "MUFU", "FFMA", "DADD", "LDG", "STG", "ATOMG",
         "REDG", "MEMBAR", "NANOSLEEP", "LDGMC", "ST.E.MC", "HMMA", "HGMMA", "STL", "LDL"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    regs = dict(re.findall(r"Function (\S+):\n\s+REG:(\d+)", res))
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (sm_100a; counts of static instructions per kernel)")
    print("# proof of the native paths: UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA),")
    print("# UBLKCP = cp.async.bulk (TMA bulk copy), UTCBAR = tcgen05.commit, USETMAXREG = setmaxnreg, LDGMC/ST.MC = multimem;")
    print("# HMMA / HGMMA (legacy mma.sync / wgmma) must be absent.\n")
    for f in re.split(r"\n\s*Function : ", sass)[1:]:
        name = f.split("\n")[0].strip()
        ops = collections.Counter()
        n = 0
        for line in f.split("\n"):
            m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if not m:
                continue
            n += 1
            op = m.group(1)
            for w in WATCH:
                if op == w or op.startswith(w + ".") or (w in ("LDGMC", "ST.E.MC") and w in op):
                    ops[w] += 1
        demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
        demangled = re.sub(r"\(anonymous namespace\)::", "", demangled)
        demangled = re.sub(r"mpv::", "", demangled)[:150]
        shown = "  ".join(f"{k}={v}" for k, v in ops.items() if v and k not in ("FFMA", "LDG", "STG") or k in ("HMMA", "HGMMA") and v)
        print(f"{demangled}\n    instr={n} regs={regs.get(name, '?')}  {shown}")


if __name__ == "__main__":
    sys.exit(main())
