"""oracle/_ref/ref_v1.{bin,json} (written by the Rust fixture generator) -> tests/golden/ref_v1.npz.

    python oracle/ref_rust/to_npz.py oracle/_ref tests/golden/ref_v1.npz
"""
import json
import os
import sys

import numpy as np


def main():
    src, dst = sys.argv[1], sys.argv[2]
    index = json.load(open(os.path.join(src, "ref_v1.json")))
    raw = open(os.path.join(src, "ref_v1.bin"), "rb").read()
    out = {}
    for name, meta in index.items():
        count = int(np.prod(meta["shape"])) if meta["shape"] else 1
        out[name] = np.frombuffer(raw, dtype="<" + meta["dtype"], count=count, offset=meta["offset"]).reshape(meta["shape"]).copy()
    np.savez_compressed(dst, **out)
    print(f"{len(out)} arrays -> {dst}")


if __name__ == "__main__":
    main()
