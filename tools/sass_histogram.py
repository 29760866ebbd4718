"""Per-kernel SASS opcode histogram of the built library (cuobjdump -sass): evidence that the hot kernels are Blackwell-native
(tcgen05 = UTC*MMA, TMEM = LDTM / STTM, TMA = UTMALDG / UTMASTG, packed FP32 = FFMA2 / FADD2 / FMUL2, ...).  No GPU needed.

    python tools/sass_histogram.py > profiles/sass_opcodes_r2.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "3d-shape-generation_b200", "libpcd_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTMACCTL", "SYNCS",
        "FFMA2", "FADD2", "FMUL2", "FMNMX3", "FMNMX", "REDUX", "ATOMS", "ATOMG", "RED", "HMMA", "LDGSTS", "UBLKCP", "ACQBULK", "F2FP", "MUFU", "BAR", "CCTL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
            cur["__total__"] += 1
    dm = demangle(list(per))
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(per)} kernels; columns = instruction counts in the SASS of each kernel (static, not dynamic)")
    print("# variant forms are folded into their base mnemonic (e.g. UTCHMMA.2CTA is listed separately as '.2CTA' in the detail column)")
    rows = []
    for fn, cnt in per.items():
        name = re.sub(r"\(.*", "", dm.get(fn, fn)).replace("pcd::", "").replace("(anonymous namespace)::", "")
        hist = collections.Counter()
        detail = collections.Counter()
        for op, n in cnt.items():
            base = op.split(".")[0]
            if base in KEYS:
                hist[base] += n
                if base.startswith("UTC") or base in ("UTMALDG", "UTMASTG", "LDTM", "STTM"):
                    detail[op] += n
        if not any(hist[k] for k in ("UTCHMMA", "UTCQMMA", "UTCOMMA", "UTCIMMA", "LDTM", "UTMALDG", "FFMA2", "REDUX", "UBLKCP")):
            continue
        rows.append((name, cnt["__total__"], hist, detail))
    rows.sort(key=lambda r: r[0])
    for name, total, hist, detail in rows:
        cols = " ".join(f"{k}={hist[k]}" for k in KEYS if hist[k])
        print(f"{name}\n    instructions={total} {cols}")
        d = " ".join(f"{k}:{v}" for k, v in sorted(detail.items()))
        if d:
            print(f"    forms: {d}")
    tot = collections.Counter()
    for _, _, hist, detail in rows:
        tot.update(detail)
    print("# totals over the listed kernels:", " ".join(f"{k}:{v}" for k, v in sorted(tot.items())))


if __name__ == "__main__":
    sys.exit(main())
