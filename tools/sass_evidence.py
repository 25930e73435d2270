"""Count the SASS mnemonics that show what each kernel of the built library is made of (TMA, mbarrier, packed FP32, FP64,
shared-memory loads, atomics, shuffles; tensor-core mnemonics for the record) -> profiles/r2_sass_evidence.md.

    python tools/sass_evidence.py [out.md]      # needs only cuobjdump (no GPU)
"""
import os, re, subprocess, sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "soundscope_b200", "libsoundscope_b200.so")
COLS = ["UTMALDG", "UBLKCP", "SYNCS", "FFMA2", "FADD2", "FFMA", "DFMA", "F2F.F64.F32", "LDS", "STG", "ATOM", "RED", "SHFL", "BAR",
        "UTMASTG", "HMMA", "UTC", "LDTM"]


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_evidence.md")
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    rows = []
    for f in re.split(r"\n\s*Function : ", sass)[1:]:
        name = f.split("\n", 1)[0].strip()
        dn = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        m = re.search(r"(k_\w+(?:<[^>]*>)?)", dn)
        short = m.group(1) if m else dn[:60]
        c, n = Counter(), 0
        for l in f.split("\n"):
            mm = re.search(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
            if not mm:
                continue
            n += 1
            op = mm.group(1)
            for col in COLS:
                base = op.split(".")[0]
                if op == col or op.startswith(col + ".") or (col == "UTC" and op.startswith("UTC")) or \
                        (col == "ATOM" and base in ("ATOMG", "ATOMS")) or (col == "RED" and base == "REDG"):
                    c[col] += 1
        rows.append((short, n, c))
    lines = ["# SASS evidence, final round-2 build of soundscope_b200/libsoundscope_b200.so (cuobjdump -sass; sm_100a)", "",
             "Produced by `python tools/sass_evidence.py`.  TMA shows up as UTMALDG (cp.async.bulk.tensor) / UBLKCP (cp.async.bulk) with",
             "SYNCS (mbarrier); packed FP32 as FFMA2 / FADD2; BAR counts include the per-pair named barriers of the fused epilogue.",
             "No UTC*MMA / LDTM / HMMA: nothing on this path is a contraction (BASELINE north_star: tensor cores not used).", "",
             "| kernel | instr | " + " | ".join(COLS) + " |", "|---|---|" + "---|" * len(COLS)]
    for short, n, c in sorted(rows):
        lines.append(f"| `{short}` | {n} | " + " | ".join(str(c[k]) for k in COLS) + " |")
    open(out, "w").write("\n".join(lines) + "\n")
    print(out, len(rows), "kernels")


if __name__ == "__main__":
    main()
