"""Integer-pipe and field-product throughput on the current GPU (writes one JSON line; bench.py embeds the same numbers)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import barretenberg_b200 as bb  # noqa: E402


def run(lib=None, iters=4096):
    lib = lib or bb.default_library()
    names = {0: "imad_lo_32", 1: "imad_wide_64acc", 2: "imad_wide_carry_chain", 3: "fq_mul", 4: "fr_mul"}
    out = {}
    for mode, name in names.items():
        ops, ms = lib.microbench(mode, iters if mode < 3 else iters // 8)
        out[name] = {"per_s": ops, "ms": ms}
    return out


if __name__ == "__main__":
    print(json.dumps(run()))
