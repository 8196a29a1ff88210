"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA code (B200_PROFILING.md): UTCHMMA (tcgen05.mma),
LDTM (tcgen05.ld), UTMALDG / UTMASTG (TMA tensor load / store), UBLKCP (bulk copy), SYNCS (mbarrier), HMMA (legacy mma.sync).

    python tools/sass_counts.py > profiles/sass_r02.txt      (runs on the CPU box: cuobjdump on the in-tree library)
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tomoencoders_b200", "libtomoenc.so")
MNEMONICS = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "STG", "LDG"]


def demangle(names):
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    return out if len(out) == len(names) else names


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    blocks = sass.split("Function : ")[1:]
    names = [b.split("\n", 1)[0].strip() for b in blocks]
    nice = demangle(names)
    rows = []
    for name, b in zip(nice, blocks):
        body = b.split("\n", 1)[1] if "\n" in b else ""
        ops = re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", body, flags=re.M)
        total = len(ops)
        cnt = {m: 0 for m in MNEMONICS}
        for op in ops:
            base = op.split(".")[0]
            if base in cnt:
                cnt[base] += 1
        short = re.sub(r"\(.*$", "", name.replace("(int)", "").replace("(bool)", "").replace("void ", "").replace("te::(anonymous namespace)::", "").replace("te::", ""))
        rows.append((short, total, cnt))
    rows.sort(key=lambda r: (-r[2]["UTCHMMA"], -r[2]["UTMALDG"], r[0]))
    print("SASS mnemonic counts per kernel of tomoencoders_b200/libtomoenc.so (cuobjdump -sass, sm_100a)")
    print("%-64s %7s " % ("kernel", "instrs") + " ".join("%8s" % m for m in MNEMONICS))
    for short, total, cnt in rows:
        if cnt["UTCHMMA"] or cnt["UTMALDG"] or cnt["UTMASTG"] or cnt["LDTM"] or cnt["HMMA"] or cnt["UBLKCP"]:
            print("%-64s %7d " % (short[:64], total) + " ".join("%8d" % cnt[m] for m in MNEMONICS))
    n_umma = sum(1 for r in rows if r[2]["UTCHMMA"])
    n_legacy = sum(1 for r in rows if r[2]["HMMA"])
    print("\n%d kernels in the library; %d issue tcgen05.mma (UTCHMMA), %d read TMEM (LDTM), %d use TMA tensor loads, %d hold legacy HMMA" % (
        len(rows), n_umma, sum(1 for r in rows if r[2]["LDTM"]), sum(1 for r in rows if r[2]["UTMALDG"]), n_legacy))


if __name__ == "__main__":
    sys.exit(main())
