#!/usr/bin/env python
"""CPU estimate of the cooperative traversal's round structure for different gate-tree shapes (tuning aid, no GPU):
rebuilds a W-wide tree over the gate boxes of every BVH item of a scene and counts, for batches of 320 rays with a
realistic bounce mix, node tasks per level -> rounds per ray-round.  usage: sim_gate_tree.py [scene] [nbatch]"""
import importlib, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
rtnw = importlib.import_module("peter-shirley-ray-tracing-the-next-week_b200")
import oracle_port as op

NONE = -2**31
scene = sys.argv[1] if len(sys.argv) > 1 else "final_northstar"
nbatch = int(sys.argv[2]) if len(sys.argv) > 2 else 24
GROUP = 320
hs = rtnw.HostScene(scene)
t = rtnw.device_tables(hs.desc_ptr)
tag = t["recs"][:, 6].copy().view(np.uint32); ival = t["recs"][:, 7].copy().view(np.int32)
ref = t["wnodes"][:, 24:28].copy().view(np.int32)
xf_ops = None

def items():
    out, i = [], 0
    while (tag[i] & 15) == 9:
        a = t["recs"][i, :3].copy().view(np.int32)
        out.append((i, int(a[0]), int(a[1]), int(a[2]), int(ival[i]) == 1, int(tag[i] >> 8)))
        i = int(a[0])
    return out

def gate_boxes(root):
    lo, hi, stack = [], [], [root]
    while stack:
        n = stack.pop()
        w = t["wnodes"][n]
        for j in range(4):
            r = int(ref[n, j])
            if r == NONE: continue
            if r >= 0: stack.append(r)
            else:
                lo.append(w[[j, 4 + j, 8 + j]]); hi.append(w[[12 + j, 16 + j, 20 + j]])
    return np.array(lo, np.float64), np.array(hi, np.float64)

def half_area(lo, hi):
    d = hi - lo
    return d[..., 0] * d[..., 1] + d[..., 1] * d[..., 2] + d[..., 2] * d[..., 0]

def build_binary(lo, hi, ids, margin_div=8):
    """returns nested tuple tree: ('leaf', gate) or ('node', lo, hi, left, right)"""
    if len(ids) == 1:
        return ("leaf", ids[0], lo[ids[0]], hi[ids[0]])
    best = (1e300, 0, len(ids) // 2)
    for axis in range(3):
        order = ids[np.argsort(lo[ids, axis] + hi[ids, axis], kind="stable")]
        l_lo = np.minimum.accumulate(lo[order], 0); l_hi = np.maximum.accumulate(hi[order], 0)
        r_lo = np.minimum.accumulate(lo[order][::-1], 0)[::-1]; r_hi = np.maximum.accumulate(hi[order][::-1], 0)[::-1]
        n = len(ids); margin = n // margin_div
        for q in range(n - 1):
            if q + 1 < margin or n - q - 1 < margin: continue
            c = half_area(l_lo[q], l_hi[q]) * (q + 1) + half_area(r_lo[q + 1], r_hi[q + 1]) * (n - q - 1)
            if c < best[0]: best = (c, axis, q + 1)
    _, axis, split = best
    order = ids[np.argsort(lo[ids, axis] + hi[ids, axis], kind="stable")]
    L = build_binary(lo, hi, order[:split], margin_div); R = build_binary(lo, hi, order[split:], margin_div)
    return ("node", np.minimum(L[2] if L[0] == "leaf" else L[1], R[2] if R[0] == "leaf" else R[1]),
            np.maximum(L[3] if L[0] == "leaf" else L[2], R[3] if R[0] == "leaf" else R[2]), L, R)

def box_of(n):
    return (n[2], n[3]) if n[0] == "leaf" else (n[1], n[2])

def collapse(bt, W):
    """W-wide nodes: list of dict(lo[k,3], hi[k,3], child[k] = ('g', gate) | ('n', index))"""
    nodes = []
    def emit(n):
        kids = [n] if n[0] == "leaf" else [n[3], n[4]]
        while len(kids) < W:
            cand = [(half_area(*box_of(k)), i) for i, k in enumerate(kids) if k[0] == "node"]
            if not cand: break
            _, i = max(cand)
            k = kids[i]; kids[i] = k[3]; kids.append(k[4])
        me = len(nodes); nodes.append(None)
        lo = np.array([box_of(k)[0] for k in kids]); hi = np.array([box_of(k)[1] for k in kids])
        ch = [("g", k[1]) if k[0] == "leaf" else ("n", emit(k)) for k in kids]
        nodes[me] = dict(lo=lo, hi=hi, child=ch)
        return me
    root = emit(bt)
    return nodes, root

def slab(lo, hi, o, inv, tmin, tmax):
    t0 = (lo - o) * inv; t1 = (hi - o) * inv
    near = np.minimum(t0, t1); far = np.maximum(t0, t1)
    a = np.maximum(np.nanmax(near, -1), tmin); b = np.minimum(np.nanmin(far, -1), tmax)
    return ~(b <= a)

def traverse(nodes, root, O, D, tmax, lanes_per_node):
    """per level: number of node tasks (over all rays of the batch); returns (levels list, gates passed, box tests)"""
    inv = 1.0 / D
    levels, gates, boxes = [], 0, 0
    cur = [(r, root) for r in range(len(O))]
    while cur:
        levels.append(len(cur))
        nxt = []
        for r, n in cur:
            nd = nodes[n]
            p = slab(nd["lo"], nd["hi"], O[r], inv[r], 0.001, tmax[r])
            boxes += len(p)
            for ok, c in zip(p, nd["child"]):
                if not ok: continue
                if c[0] == "g": gates += 1
                else: nxt.append((r, c[1]))
        cur = nxt
    return levels, gates, boxes

# ---- rays with a bounce mix: camera rays, then cosine-ish scattered rays from the hit points (oracle closest hits)
rng = np.random.default_rng(1)
nx = ny = 1000
cam = hs.camera(nx, ny)
def xform(o, d, chain):
    ops = hs.desc.xforms
    if chain == 0: return o, d
    n = ops[chain].kind >> 8
    o = o.copy(); d = d.copy()
    for k in range(n):
        opk = ops[chain + k]
        if (opk.kind & 255) == 1: o = o - np.array([opk.a, opk.b, opk.c])
        else:
            s, c = opk.a, opk.b
            o = np.stack([c * o[:, 0] - s * o[:, 2], o[:, 1], s * o[:, 0] + c * o[:, 2]], 1)
            d = np.stack([c * d[:, 0] - s * d[:, 2], d[:, 1], s * d[:, 0] + c * d[:, 2]], 1)
    return o, d

its = [it for it in items() if it[4]]
trees = {}
for W in (4, 8, 16):
    trees[W] = []
    for (i, nxt, root, depth, _, chain) in its:
        lo, hi = gate_boxes(root)
        bt = build_binary(lo, hi, np.arange(len(lo)))
        trees[W].append(collapse(bt, W))
for W in trees:
    print(f"W={W}: nodes per item {[len(n) for n, _ in trees[W]]}")

tot = {W: dict(rounds=0.0, tasks=0, boxes=0, lanes=0, levels=0) for W in trees}
nrays = 0
for b in range(nbatch):
    ij = np.stack([rng.integers(0, nx, GROUP), rng.integers(0, ny, GROUP)], 1)
    rays = rtnw_rays = op.camera_rays(rtnw, cam, nx, ny, ij, np.full(GROUP, b, np.int32), seed=3)
    alive = np.ones(GROUP, bool)
    for bounce in range(4):
        hits = op.trace(rtnw, hs.desc_ptr, rays, seed=3)
        O = rays["origin"].astype(np.float64); D = rays["direction"].astype(np.float64)
        act = np.nonzero(alive)[0]
        if len(act) == 0: break
        nrays += len(act)
        # tmax0 approx: FLT_MAX for the first BVH item, hit t for later ones (narrowing) - use the final hit t as an upper-bound proxy for item > 0
        for W in trees:
            rounds = 0
            for k, ((nodes, root), it) in enumerate(zip(trees[W], its)):
                o, d = xform(O[act], D[act], it[5])
                tmax = np.full(len(act), 3.4e38) if k == 0 else np.where(hits["prim_id"][act] >= 0, hits["t"][act].astype(np.float64) * 1.0000001, 3.4e38)
                lv, g, bx = traverse(nodes, root, o, d, tmax, 1)
                lpn = max(1, W // 4)
                rounds += sum(int(np.ceil(x * lpn / GROUP)) for x in lv) + 1
                tot[W]["tasks"] += sum(lv); tot[W]["boxes"] += bx; tot[W]["lanes"] += sum(lv) * lpn; tot[W]["levels"] += len(lv)
            tot[W]["rounds"] += rounds
        # next bounce: scatter from hit points into the normal's hemisphere (lambertian-like); misses die
        hit = hits["prim_id"] >= 0
        alive = alive & hit & (rng.random(GROUP) < 0.85)
        v = rng.normal(size=(GROUP, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
        nd = hits["normal"] + v
        rays = rays.copy()
        rays["origin"] = hits["p"]; rays["direction"] = nd.astype(np.float32)
nb = nbatch
for W in tot:
    T = tot[W]
    print(f"W={W}: rounds/ray-round(bounce) {T['rounds'] / (nb * 4):.2f}  node tasks/ray {T['tasks'] / nrays:.2f}  lane-tasks/ray {T['lanes'] / nrays:.2f}  box tests/ray {T['boxes'] / nrays:.1f} levels/item {T['levels'] / (nb * 4 * len(its)):.2f}")
