"""profiles/sass_rNN.txt: per-kernel counts of the SASS mnemonics that prove (or disprove) a Blackwell-native kernel.
UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tile load, UTCBAR = tcgen05.commit,
HMMA (not preceded by UTC) = legacy mma.sync — must be 0.   Usage: python tools/sass_table.py > profiles/sass_r02.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "dalle_mtf_b200", "libdalle_b200.so")
PATS = [("UTCHMMA", r"UTCHMMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"UTMALDG"),
        ("UTMAPF", r"UTMAPF"), ("UTCBAR", r"UTCBAR"), ("SYNCS", r"\bSYNCS"), ("HMMA", r"(?<!UTC)HMMA"),
        ("MUFU.EX2", r"MUFU\.EX2"), ("REDG", r"\bREDG"), ("ATOMG", r"\bATOMG"), ("LDG", r"\bLDG"), ("STG", r"\bSTG")]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    counts, kern = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            kern = m.group(1)
            counts[kern] = collections.Counter()
            continue
        if kern and "/*" in line:
            for name, pat in PATS:
                if re.search(pat, line):
                    counts[kern][name] += 1
    names = subprocess.run(["c++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
    print("# SASS evidence: instruction counts per kernel of dalle_mtf_b200/libdalle_b200.so (cuobjdump -sass)")
    print("# UTCHMMA = tcgen05.mma | LDTM / STTM = tcgen05.ld / st | UTMALDG = TMA load | UTCBAR = tcgen05.commit |"
          " HMMA = legacy mma.sync (0 everywhere)")
    print(f"{'kernel':64s} " + " ".join(f"{n:>8s}" for n, _ in PATS))
    tot = collections.Counter()
    for (k, c), nm in zip(counts.items(), names):
        nm = re.sub(r"\(.*", "", nm).replace("db200::", "").replace("(anonymous namespace)::", "").replace("void ", "")
        print(f"{nm[:64]:64s} " + " ".join(f"{c[n]:8d}" for n, _ in PATS))
        tot.update(c)
    print(f"{'TOTAL':64s} " + " ".join(f"{tot[n]:8d}" for n, _ in PATS))


if __name__ == "__main__":
    main()
