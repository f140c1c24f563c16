#!/usr/bin/env python
"""tests/golden/config_scalars.json: physics scalars of the five BASELINE configs computed from matrices
produced by the COMPILED REFERENCE (oracle/_ref) and assembled by scipy, exactly as reference scripts do."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_loop  # noqa: E402
from tests import configs  # noqa: E402


def main():
    out = {}
    for name in configs.NAMES:
        cs = configs.build(name)
        keys = configs.what(name)
        outs = [ref_loop.run(c, what=[k for k in keys if not (c["kind"] == "beamc" and k == "KGs")]) for c in cs]
        mats = {k: configs.assemble_scipy(cs, outs, k) for k in keys}
        out[name] = configs.scalars(name, cs, mats)
        print(name, out[name])
    with open(os.path.join(HERE, "config_scalars.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
