#!/usr/bin/env python
"""Shape sweep of the attention op (kernel without row maxima + its classic twin) against fp32 torch:
token counts around every tile boundary, odd numbers of (frame, head)s, 1 / 6 / 12 heads.  Prints one line per
failing case and a summary; exit code 1 if any case fails."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, ROOT)
import gpu_check  # noqa: E402


def main():
    import io
    import contextlib
    import json
    bad = 0
    n = 0
    for N in (1, 2, 63, 65, 127, 128, 129, 255, 256, 257, 383, 384, 385, 511, 513, 640, 901, 1025, 1153):
        for B, H in ((1, 1), (1, 6), (3, 1), (2, 12), (3, 6)):
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                ok = gpu_check.check_attention(f"attn_B{B}_N{N}_H{H}", B, N, H)
            n += 1
            if not ok:
                bad += 1
                print(buf.getvalue().strip()[:400])
    print(json.dumps({"cases": n, "failed": bad, "unshifted": os.environ.get("DINOSEG_ATTN_UNSHIFTED", "1")}))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
