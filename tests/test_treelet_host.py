"""Host-side check of the BVH builder's binary-tree stages: ospray_b200/csrc/cuda/treelet.h (Karras radix tree node,
7-leaf treelet restructuring with the surface-area cost) compiles for the host too; tests/native/treelet_check.cpp runs the
passes in the GPU kernel's bottom-up order and validates the tree after each (every leaf reachable once, parents,
exact boxes, leaf counts) and that the SAH cost goes down.  Stands in for Embree's SAH builder behind rtcCommitScene
(modules/cpu/common/Group.cpp:129-142), which is not vendored."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("treelet") / "treelet_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", os.path.join(ROOT, "tests", "native", "treelet_check.cpp"), "-o", exe])
    return exe


@pytest.mark.parametrize("n,dist", [(2, 0), (7, 1), (8, 0), (9, 2), (33, 1), (1000, 2), (20000, 0), (20000, 1), (20000, 2)])
def test_treelet_passes_keep_the_tree_valid_and_lower_the_cost(checker, n, dist):
    p = subprocess.run([checker, str(n), str(dist), "3"], capture_output=True, text=True)
    res = json.loads(p.stdout)
    assert p.returncode == 0 and res["valid"] == [0, 0, 0, 0], res
    cost = res["cost"]
    assert all(b <= a * (1 + 1e-6) for a, b in zip(cost, cost[1:])), cost   # a pass never raises the cost
    if n >= 20000:
        assert cost[-1] < (0.6 if dist == 1 else 0.95) * cost[0], cost       # big-among-small primitives gain most
        assert res["rewritten"][0] > res["rewritten"][-1] > 0


def test_traversal_stack_need_of_large_trees(checker):
    """the builder reports the deepest stack a depth-first walk of the collapsed 4-wide hierarchy can need (collapseKernel's
    `pushed`; restated in tests/native/treelet_check.cpp); the device refuses hierarchies beyond OSPCUDA_TRAVERSAL_STACK = 48.
    A jittered lattice of boxes (the 10 M-triangle bench mesh in small) stays well inside; the refinement does not deepen it."""
    for passes in (0, 2):
        res = json.loads(subprocess.run([checker, "300000", "3", str(passes)], capture_output=True, text=True).stdout)
        assert res["valid"] == [0] * (passes + 1)
        assert res["wide_depth"] <= 14 and res["stack_need"] <= 40, res
