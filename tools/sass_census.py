#!/usr/bin/env python
"""Per-kernel SASS census of libmogstn.so (cuobjdump -sass, sm_100a): instruction totals and the mnemonics that matter.

    python tools/sass_census.py > profiles/r02_sass_mnemonics.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "mog_asr_b200", "libmogstn.so")
KEYS = ["LDG", "STG", "LDS", "STS", "UTMALDG", "UBLKCP", "SYNCS", "SHFL", "REDG", "ATOMG", "FMUL", "FADD", "FFMA",
        "IMAD", "BAR", "LDL", "STL", "HMMA", "UTCMMA"]

HEADER = """# cuobjdump -sass mog_asr_b200/libmogstn.so (sm_100a), per kernel: instruction count and the mnemonics that matter
# UTMALDG = cp.async.bulk.tensor (TMA tensor load: the CTA-per-image and per-warp-ring backward kernels); SYNCS = mbarrier
# arrive/expect_tx/try_wait; UBLKCP = cp.async.bulk shared->global (bulk-copy engine zero fill); REDG = red.global.add.f32
# (general-affine backward cold path, ASR column sums, second-layer weight gradients of the fused heads); no HMMA/UTCMMA
# anywhere: nothing on this path is a dense contraction (SURVEY section 8)
"""


def main():
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    out = [HEADER]
    name, counts = None, None

    def flush():
        if name:
            parts = [f"total={counts['total']}"] + [f"{k}={counts[k]}" for k in KEYS if counts[k]]
            out.append(name + "\n    " + "  ".join(parts))
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            flush()
            name, counts = m.group(1), collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and name:
            counts["total"] += 1
            op = m.group(1)
            for k in KEYS:
                if op == k or op.startswith(k + "."):
                    counts[k] += 1
    flush()
    sys.stdout.write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
