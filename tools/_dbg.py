"""debugging aid (GPU box): the crowded-cell search case, missing / extra pairs against the oracle"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import apbf_b200 as gpu
from apbf_b200 import scenes
from oracle import oracle as orc
from test_gpu_parity import oracle_state

rng = np.random.default_rng(17)
sc = scenes.uniform_block(8, jitter=0.0, res_log2=4)
n = sc.n
pos = sc.arrays["position"]
pos[: n // 2, :3] = pos[0, :3]
pos[n // 2: 3 * n // 4, :3] = pos[0, :3] + rng.integers(-40000, 40000, (n // 4, 3)).astype(np.int32)
sc.arrays["pos_backup"][:] = pos
sc.arrays["kernel_width"] = (sc.arrays["kernel_width"] * rng.choice([0.0, 0.3, 1.0, 2.5], n)).astype(np.float32)
s = orc.default_settings()
cap = n * n
st = oracle_state(orc, sc)
ep = orc.green_apply(st, s, 3, 1.0, sc.min_pos, sc.max_pos, sc.res_log2, cap)
ctx = gpu.Context()
L = gpu.ParticleLists(ctx, sc.arrays, neighbor_capacity=cap)
gpu.neighborhood_green(ctx).set_data(L).set_range_scale(1.0).set_position_range(sc.min_pos, sc.max_pos, sc.res_log2).apply()
gp = L.read_pairs()
print("res", sc.res_log2, "min", sc.min_pos, "max", sc.max_pos, "oracle", len(ep), "gpu", len(gp))
es = set(map(tuple, ep.tolist())); gs = set(map(tuple, gp.tolist()))
miss = sorted(es - gs); extra = sorted(gs - es)
print("missing", len(miss), "extra", len(extra))
p = st.position[:, :3].astype(np.float64) / 262144.0
for a, b in miss[:20]:
    d = np.linalg.norm(p[a] - p[b])
    print("miss", a, b, "kw_a", st.kernel_width[a], "kw_b", st.kernel_width[b], "d", d, "pa", p[a], "pb", p[b])
for a, b in extra[:10]:
    print("extra", a, b)
