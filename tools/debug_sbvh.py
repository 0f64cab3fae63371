"""Where a GPU high-quality build differs from the oracle's: the binary trees are walked together from the root (children
matched by the set of primitives below them); the first node whose two sides differ is printed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.setrecursionlimit(100000)
import numpy as np
from hiprt_b200 import api, scenes
from oracle import binding as orc
from tests.test_gpu_parity import to_dev
from tests.test_oracle_sbvh import CASES
ctx = api.Context(0)

def tree(nodes, leaf_prim):
    """per node: (child0, child1) and the leaf sets (as sorted tuples of leaf-record primitives) by DFS"""
    memo = {}
    def below(c):
        if (c & 7) == 5:
            k = c >> 4
            if k not in memo:
                memo[k] = tuple(sorted(below(int(nodes["child0"][k])) + below(int(nodes["child1"][k]))))
            return memo[k]
        return (int(leaf_prim[c >> 3]),)
    below(5)
    return memo, below

for name in sys.argv[1:] or ["sphere"]:
    for spatial in (False, True):
        verts, idx = CASES[name]()
        flags = api.HIGH_QUALITY_BUILD | (0 if spatial else 4)
        geom = ctx.build(to_dev(ctx, verts), None if idx is None else to_dev(ctx, idx), flags=flags)
        og = orc.Geometry(verts, idx, high_quality=True, spatial_splits=spatial)
        n = og.prim_count
        L = api.high_quality_temp_layout(geom.ctx_prim_max(), geom.ctx_prim_max() if idx is not None else 0, spatial)
        ctx.torch.cuda.synchronize()
        st = geom.temp[64:128].cpu().numpy().view(np.uint32)
        leaves_n, nodes_n = int(st[3]), int(st[4])
        gn = geom.temp[L["nodes"]: L["nodes"] + nodes_n * 32].cpu().numpy().view(api.SCRATCH_DTYPE)
        gl = geom.temp[L["leaves"]: L["leaves"] + leaves_n * 32].cpu().numpy().view(np.uint32).reshape(-1, 8)
        gl_box = gl[:, :6].view(np.float32)
        pairs = og.pairs
        first_to_prim = {int(a): p for p, (a, b) in enumerate(pairs)}
        gl_prim = np.array([first_to_prim[int(w)] for w in gl[:, 6]])
        on, ol_prim, ol_box = og.scratch, og.leaf_record_prims, og.leaf_record_boxes
        print(name, "spatial", spatial, "gpu nodes/leaves", nodes_n, leaves_n, "oracle", len(on), len(ol_prim), "stats", og.sbvh_stats)
        boxes, leaves = geom.nodes()
        cb, cl = orc.canonicalize(boxes.view(orc.BOX4_DTYPE), leaves.view(orc.TRIPAIR_DTYPE))
        ob = og.box_nodes
        m = min(len(cb), len(ob))
        a, b = cb.view(np.uint32).reshape(-1, 32)[:m], ob.view(np.uint32).reshape(-1, 32)[:m]
        bad = np.nonzero((a != b).any(1))[0]
        print(" wide nodes", len(cb), len(ob), "differing", len(bad))
        for k in bad[:3]:
            print(" node", k, "words", np.nonzero(a[k] != b[k])[0])
            print("  gpu child", cb["child"][k], "\n", cb["box"][k].reshape(4, 6))
            print("  orc child", ob["child"][k], "\n", ob["box"][k].reshape(4, 6))
        continue
        gm, gbelow = tree(gn, gl_prim)
        om, obelow = tree(on, ol_prim)
        shown = 0
        stack = [(5, 5, 0)]
        while stack and shown < 3:
            g, o, depth = stack.pop()
            if (g & 7) != 5 or (o & 7) != 5:
                continue
            gk, ok = g >> 4, o >> 4
            gc = [int(gn["child0"][gk]), int(gn["child1"][gk])]
            oc = [int(on["child0"][ok]), int(on["child1"][ok])]
            gs, os_ = [gbelow(c) for c in gc], [obelow(c) for c in oc]
            if gs == os_:
                stack += [(gc[0], oc[0], depth + 1), (gc[1], oc[1], depth + 1)]
                continue
            shown += 1
            print(" depth", depth, "node box gpu", gn["box"][gk].reshape(-1), "orc", on["box"][ok].reshape(-1))
            print("  gpu sides", [len(x) for x in gs], gs[0][:8], gs[1][:8])
            print("  orc sides", [len(x) for x in os_], os_[0][:8], os_[1][:8])
            prims = sorted(set(gs[0] + gs[1]))
            if len(prims) <= 6:
                for p in prims:
                    print("   prim", p, "pair", pairs[p], "box", og.prim_boxes[p].reshape(-1))
