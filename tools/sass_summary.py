"""Per-kernel counts of the Blackwell-native SASS mnemonics in libcgan3d.so (profiles/rNN_sass_summary.txt).

    python tools/sass_summary.py > profiles/r02_sass_summary.txt

tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM / STTM, TMA -> UTMALDG / UTMASTG / UBLKCP, cp.async -> LDGSTS
(B200_PROFILING.md: "What proves a Blackwell-native kernel")."""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "contrast_gan_3d_b200" / "libcgan3d.so"
PAT = {"UTCHMMA": r"\bUTC\w*MMA\b", "UTCHMMA.2CTA": r"\bUTC\w*MMA\.2CTA", "LDTM": r"\bLDTM\b", "STTM": r"\bSTTM\b", "UTMALDG": r"\bUTMALDG\b",
       "UBLKCP": r"\bUBLKCP\b", "UTCBAR": r"\bUTCBAR\b", "SYNCS": r"\bSYNCS\b", "LDGSTS": r"\bLDGSTS\b", "HMMA": r"\bHMMA\b"}


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
    dem = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.splitlines()
    names = iter(dem)
    counts, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = next(names, m.group(1))
            cur = re.sub(r"\(.*", "", cur).replace("void ", "")
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for k, p in PAT.items():
            if re.search(p, line):
                counts[cur][k] += 1
    tot = collections.Counter()
    print(f"# {LIB.name}: cuobjdump -sass, mnemonic counts per kernel (sm_100a)")
    print(f"{'kernel':78s} " + " ".join(f"{k:>12s}" for k in PAT))
    for name, c in counts.items():
        tot.update(c)
        if sum(c[k] for k in PAT if k not in ("SYNCS",)) == 0:
            continue
        print(f"{name[:78]:78s} " + " ".join(f"{c[k]:12d}" for k in PAT))
    print(f"{'TOTAL (' + str(len(counts)) + ' kernels)':78s} " + " ".join(f"{tot[k]:12d}" for k in PAT))


if __name__ == "__main__":
    sys.exit(main())
