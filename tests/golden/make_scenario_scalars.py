#!/usr/bin/env python
"""tests/golden/scenario_scalars.json: the physics scalars of tests/scenarios.py computed from matrices produced by the
COMPILED REFERENCE (oracle/_ref element loop + scipy coo -> csr), exactly as reference scripts do.  Build container only."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_loop  # noqa: E402
from tests import scenarios  # noqa: E402


def main():
    assert ref_loop.available(), "build oracle/_ref first: python oracle/build_ref.py"
    out = {}
    ev = scenarios.evaluate_with(ref_loop.run)
    for name in scenarios.NAMES:
        out[name] = scenarios.SCENARIOS[name](ev)
        print(name, out[name])
    with open(os.path.join(HERE, "scenario_scalars.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
