"""Per-kernel SASS mnemonic census of the built library -> profiles/r2_sass_census.md.

    python tools/sass_census.py [out.md]

Proof that the tensor-core kernels are tcgen05 / TMEM / TMA code (UTCHMMA, LDTM, UTMALDG, UTMASTG) and that nothing in the
library uses the legacy warp-level mma path (HMMA).  Needs cuobjdump and c++filt; no GPU."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "msfwsi_b200", "lib", "libmsfwsi_b200.so")
PAT = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "MUFU.EX2", "HMMA", "ATOMG", "REDG"]
HEAD = """# SASS mnemonic census per kernel (round 2)

`cuobjdump -sass msfwsi_b200/lib/libmsfwsi_b200.so` (sm_100a), instruction counts per kernel; regenerate with `tools/sass_census.py`.

UTCHMMA = tcgen05.mma kind::f16; LDTM / STTM = tcgen05.ld / st (TMEM <-> registers); UTMALDG / UTMASTG = TMA tensor load / store;
UTCBAR = tcgen05.commit -> mbarrier; SYNCS = mbarrier arrive / try_wait; HMMA = legacy mma.sync (none anywhere: no kernel of this
library uses the warp-level tensor-core path); ATOMG / REDG = global atomics (split-K arrival counters, exchange flags, ticket locks).
"""


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_census.md")
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)[1:]
    names = [f.split("\n", 1)[0].strip() for f in funcs]
    dems = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.strip().split("\n")
    rows = []
    for f, dem in zip(funcs, dems):
        cnt = {p: len(re.findall(r"\b" + re.escape(p), f)) for p in PAT}
        n = len(re.findall(r"/\*[0-9a-f]{4}\*/", f))
        d = re.sub(r"\(.*$", "", dem.replace("(anonymous namespace)::", "").replace("void ", ""))
        rows.append((d, n, cnt))
    lines = [HEAD, "| kernel | SASS instr | " + " | ".join(PAT) + " |", "|---|---:|" + "---:|" * len(PAT)]
    for d, n, cnt in sorted(rows, key=lambda r: (-(r[2]["UTCHMMA"] > 0), r[0])):
        lines.append(f"| `{d}` | {n} | " + " | ".join(str(cnt[p]) if cnt[p] else "" for p in PAT) + " |")
    open(out, "w").write("\n".join(lines) + "\n")
    print(f"{len(rows)} kernels -> {out}")


if __name__ == "__main__":
    main()
