"""CPU: host-side logic of bench.py — the strip partition every rank builds for itself (weak and strong scaling) is a
disjoint cover of nodes and elements, and both arms name the same workload."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("weak", [True, False])
def test_plate_shards_cover_the_mesh_once(world, weak):
    import bench
    from pyfe3d_b200 import meshes
    side = 16
    nx, ny = (side * world, side) if weak else (side, side)
    a = float(world) if weak else 1.0
    nny = ny + 1
    owned_nodes, owned_elems, seen = 0, 0, np.zeros((nx + 1) * nny, int)
    for rank in range(world):
        case = bench.plate_shard(meshes, nx, ny, a, rank, world)
        lo, hi = case["owned_nodes"]
        off = case["node_offset"]
        conn = case["conn"]
        assert conn.min() >= 0 and conn.max() < case["ndof"] // 6
        # every element this rank evaluates touches an owned node; every element touching an owned node is present
        touches = ((conn >= lo) & (conn < hi)).any(1)
        assert touches.all()
        owned_nodes += hi - lo
        owned_elems += case["owned_elements"]
        assert case["owned_elements"] == int(((conn[:, 0] >= lo) & (conn[:, 0] < hi)).sum())
        seen[off + lo:off + hi] += 1
        # the rank's coordinates are the global ones of its node columns
        full = meshes.plate_quad4(nx, ny, a=a, b=1.0)
        g = np.asarray(full["x"]).reshape(-1, 3)[off:off + case["ndof"] // 6]
        assert np.allclose(np.asarray(case["x"]).reshape(-1, 3), g, rtol=0, atol=1e-15)
    assert owned_nodes == (nx + 1) * nny and owned_elems == nx * ny
    assert (seen == 1).all()


def test_reference_arm_line_names_the_same_workload():
    import bench
    from oracle import ref_loop
    if not ref_loop.available():
        pytest.skip("oracle/_ref not built")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-side", "16"], capture_output=True, text=True, timeout=600)
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
    assert line["config"] == {"workload": bench.workload_name(2000), "l2": bench.L2_NOTE}
    assert line["cpu_baseline"]["kind"] == "reference" and line["e2e"]["value"] == line["value"]
