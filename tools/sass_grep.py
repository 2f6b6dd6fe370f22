#!/usr/bin/env python
"""SASS mnemonic counts per kernel of libndmps_sm100.so -> profiles/rNN_sass_grep.txt (no GPU needed).

    python tools/sass_grep.py [round]           # default 02
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "img-compression-mps_b200" / "imgcompressionmps" / "libndmps_sm100.so"
WANT = ("UTCHMMA", "UTCIMMA", "UTCQMMA", "LDTM", "UTMALDG", "UTCBAR", "UTCATOMSWS", "DMMA", "UBLKCP", "SYNCS", "UCGABAR", "LDGSTS")
HEADER = """# SASS mnemonic counts per kernel of libndmps_sm100.so (cuobjdump -sass), round {rnd}; tools/sass_grep.py
# tcgen05.mma -> UTC*MMA (UTCHMMA: kind::f16, UTCIMMA: kind::i8), tcgen05.ld -> LDTM, cp.async.bulk.tensor -> UTMALDG,
# tcgen05.commit -> UTCBAR, tcgen05.alloc -> UTCATOMSWS, mma.sync f64 -> DMMA, cp.async.bulk -> UBLKCP, mbarrier -> SYNCS,
# barrier.cluster -> UCGABAR, cp.async -> LDGSTS
"""


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else "02"
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    counts, order, cur, idx = {}, [], None, 0
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = re.sub(r"\(.*", "", names[idx])
            idx += 1
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            for w in WANT:
                if op.startswith(w):
                    counts[cur][w] += 1
    out = [HEADER.format(rnd=rnd)]
    total = collections.Counter()
    for k in order:
        if counts[k]:
            out.append(f"{k:<100} " + " ".join(f"{w}={c}" for w, c in sorted(counts[k].items())))
            total.update(counts[k])
    out.append("")
    out.append("TOTAL " + " ".join(f"{w}={c}" for w, c in sorted(total.items())))
    path = ROOT / "profiles" / f"r{rnd}_sass_grep.txt"
    path.write_text("\n".join(out) + "\n")
    print(path, "TOTAL", dict(total))


if __name__ == "__main__":
    main()
