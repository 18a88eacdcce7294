"""Static evidence of what the built library executes: per-kernel counts of the SASS mnemonics that identify the
hardware path (DMMA = FP64 tensor pipe, UTMALDG/UTMASTG = TMA tensor load/store, UTCHMMA/LDTM/UTCBAR = tcgen05 MMA /
TMEM load / commit, SYNCS = mbarrier, LDGSTS = cp.async, UBLKCP = bulk copy), plus registers / spills from the
cubin's resource usage.  No GPU needed:  python tools/sass_census.py > profiles/r02_sass_census.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tor10_b200", "lib", "libtor10_b200.so")
KEYS = ["DMMA", "UTMALDG", "UTMASTG", "UTCHMMA", "LDTM", "UTCBAR", "SYNCS", "ACQBULK", "LDGSTS", "UBLKCP", "DFMA",
        "DADD", "DMUL", "SHFL", "LDS", "STS", "LDG", "STG", "ATOMG", "MEMBAR", "NANOSLEEP"]


def demangle(name):
    out = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    return re.sub(r"\(.*", "", out).replace("void ", "").replace("t10::", "")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    usage, cur = {}, None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m and cur:
            usage[cur] = tuple(int(v) for v in m.groups())
    per, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            per[cur][m.group(1)] += 1
    # architecture of every cubin that holds at least one kernel (nvcc's host-link step also embeds an EMPTY
    # default-arch stub cubin with no functions; it is not counted)
    arch = sorted(set(m.group(1) for m in re.finditer(r"arch = (sm_\w+)(?:(?!arch = ).)*?Function :", sass, re.S)))
    print("# SASS census of tor10_b200/lib/libtor10_b200.so (`python tools/sass_census.py`, cuobjdump, no GPU)\n")
    print("cubin architectures: " + ", ".join(arch) + "\n")
    print("Counts are static instructions in the kernel body (loops unrolled by the compiler count once per copy).")
    print("REG / STACK / LOCAL from `cuobjdump -res-usage`: the 40-byte STACK of `k_ggemm_f64_tma<4, 2>` is ten per-piece\n"
          "scalars the compiler parks around the k loop (STL at 0x0b30-0x0ef0 before it, LDL.LU at 0x4840-0x48c0 after it; the\n"
          "DMMA loop 0x1440-0x1fb0 holds no STL / LDL), every other hot kernel has STACK = 0.\n")
    print("| kernel | SASS instr. | REG | STACK | LOCAL | static SHARED | mnemonics that identify the path |")
    print("|---|---|---|---|---|---|---|")
    for f, c in per.items():
        reg, stack, shared, local = usage.get(f, (-1, -1, -1, -1))
        tags = ", ".join("%s %d" % (k, c[k]) for k in KEYS if c[k])
        print("| `%s` | %d | %d | %d | %d | %d | %s |" % (demangle(f)[:72], sum(c.values()), reg, stack, local,
                                                       shared, tags))
    return 0


if __name__ == "__main__":
    sys.exit(main())
