"""SASS mnemonic histogram of every kernel in libpackppi_b200.so (`cuobjdump -sass`): the Blackwell-path mnemonics
(UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UBLKCP = TMA, SYNCS = mbarrier) and the arithmetic /
memory mix per kernel.  Writes profiles/<tag>_sass_histogram.txt.  Runs without a GPU."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
LIB = os.path.join(ROOT, "packppi_b200", "csrc", "libpackppi_b200.so")
KEY = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "FFMA", "FFMA2", "FMUL",
       "FADD", "MUFU", "LDG", "STG", "LDS", "STS", "LDGSTS", "SHFL", "BAR", "ATOM", "RED", "LDL", "STL")


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    text = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in text.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    names = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    out = [f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: instructions per kernel (static counts), sm_100a", ""]
    for (mangled, cnt), name in zip(kernels.items(), names):
        total = sum(cnt.values())
        name = re.sub(r"\(.*", "", name)
        picked = "  ".join(f"{k} {cnt[k]}" for k in KEY if cnt.get(k))
        out.append(f"{name}\n    total {total}  |  {picked}")
    path = os.path.join(ROOT, "profiles", f"{tag}_sass_histogram.txt")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
    print(f"wrote {path}: {len(kernels)} kernels")


if __name__ == "__main__":
    main()
