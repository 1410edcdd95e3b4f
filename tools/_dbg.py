import sys, os, numpy as np
sys.path.insert(0, '.')
import apbf_b200 as gpu
from apbf_b200 import scenes
from oracle import oracle as orc
orc.build()
sc = scenes.uniform_block(96, jitter=0.2, dims=2, shuffle=True)
print(sc.min_pos, sc.max_pos, sc.res_log2, sc.n)
s = orc.default_settings(); cap = sc.n*80
st = orc.State(**{k: v.copy() for k, v in sc.arrays.items()})
ep = orc.green_apply(st, s, 2, 1.0, sc.min_pos, sc.max_pos, sc.res_log2, cap)
ctx = gpu.Context(dims=2)
L = gpu.ParticleLists(ctx, sc.arrays, neighbor_capacity=cap)
_ = gpu.neighborhood_green(ctx).set_data(L).set_range_scale(1.0).set_position_range(sc.min_pos, sc.max_pos, sc.res_log2).apply(debug=True)
aux = _
p = L.read_pairs()
print(len(p), len(ep))
ce = np.bincount(ep[:,0], minlength=sc.n); cg = np.bincount(p[:,0], minlength=sc.n)
off = aux["pair_offsets"]; cc = np.diff(off.astype(np.int64))
print("count-pass counts vs oracle: mismatches", int((cc != ce).sum()), "total", off[-1], "first bad", np.nonzero(cc != ce)[0][:10], cc[np.nonzero(cc != ce)[0][:10]], ce[np.nonzero(cc != ce)[0][:10]])
bad = np.nonzero(ce != cg)[0]
print("bad ids", len(bad), bad[:20], ce[bad[:20]], cg[bad[:20]])
pos = L.read("position")
print(pos[bad[:10]] / 262144.0)
pf = pos[:, :3] / 262144.0
for a in bad[:3]:
    e = ep[ep[:,0]==a][:,1]; g_ = p[p[:,0]==a][:,1]
    print("id", a, "exp", e, "got", g_)
    for b in sorted(set(g_.tolist()) ^ set(e.tolist())):
        print("   diff b", b, "dist", np.linalg.norm(pf[a]-pf[b]), "pos", pf[b])
kw = L.read("kernel_width"); print("kw", kw[:5], kw.min(), kw.max())
