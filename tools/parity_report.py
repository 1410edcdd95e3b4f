#!/usr/bin/env python
"""Parity report (runs on the GPU box): CUDA path through the C-ABI vs the CPU oracle.

Prints and writes gpurun_out/parity_report.json:
  * per-pass agreement of one incompressibility iteration (integer accumulators, lambda, position shifts),
  * density-error statistics |rho/rho0 - 1| over 100 substeps for both arms (north_star).
The oracle is used here as the checker only.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import apbf_b200  # noqa: E402
from apbf_b200 import scenes  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def one_iteration(sc, hk=1, gk=1, scale=1.0, cap_per=80):
    s = orc.default_settings()
    s.mHeightKernelId, s.mGradientKernelId = hk, gk
    st = orc.State(**{k: v.copy() for k, v in sc.arrays.items()})
    cap = sc.n * cap_per
    ep = orc.green_apply(st, s, sc.dims, scale, sc.min_pos, sc.max_pos, sc.res_log2, cap)
    ctx = apbf_b200.Context(dims=sc.dims)
    ctx.set_settings(apbf_b200.Settings.from_buffer_copy(bytes(s)))
    L = apbf_b200.ParticleLists(ctx, sc.arrays, neighbor_capacity=cap)
    apbf_b200.neighborhood_green(ctx).set_data(L).set_range_scale(scale).set_position_range(sc.min_pos, sc.max_pos, sc.res_log2).apply()
    pairs_equal = bool(np.array_equal(L.read_pairs(), ep))
    before = st.position.copy()
    ea = orc.incompressibility_apply(st, s, sc.dims, ep, want_aux=True)
    ga = apbf_b200.incompressibility(ctx).set_data(L).apply(debug=True)
    got = L.read("position")
    out = {"scene": sc.name, "n": sc.n, "pairs": int(len(ep)), "pairs_bit_exact": pairs_equal, "kernels": [hk, gk]}
    for k in ("density", "sq_grad_sum"):
        d = np.abs(ga[k].astype(np.int64) - ea[k].astype(np.int64))
        out[k] = {"max_abs_units": int(d.max()), "frac_equal": float((d == 0).mean()),
                  "max_rel": float((d / np.maximum(ea[k].astype(np.float64), 1)).max())}
    d = np.abs(ga["grad_sum"].astype(np.int64) - ea["grad_sum"].astype(np.int64))
    out["grad_sum"] = {"max_abs_units": int(d.max()), "frac_equal": float((d == 0).all(axis=1).mean())}
    same = (ga["density"] == ea["density"]) & (ga["sq_grad_sum"] == ea["sq_grad_sum"]) & np.all(ga["grad_sum"] == ea["grad_sum"], axis=1)
    rel = np.abs(ga["lam"] - ea["lam"]) / np.maximum(np.abs(ea["lam"]), 1e-30)
    out["lambda"] = {"max_rel_where_acc_equal": float(rel[same].max()), "max_rel_all": float(rel.max()),
                     "frac_bit_equal": float((ga["lam"] == ea["lam"]).mean()), "frac_negative": float((ea["lam"] < 0).mean())}
    se = st.position[:, :3].astype(np.int64) - before[:, :3]
    sg = got[:, :3].astype(np.int64) - before[:, :3]
    err = np.abs(sg - se)
    out["position_shift"] = {"max_shift_units": int(np.abs(se).max()), "max_err_units": int(err.max()),
                             "mean_err_units": float(err.mean()), "frac_exact": float((err == 0).all(axis=1).mean()),
                             "hist_err_units": np.bincount(err.max(axis=1).clip(0, 8), minlength=9).tolist(),
                             "max_err_rel_to_max_shift": float(err.max() / max(np.abs(se).max(), 1))}
    return out


def density_stats(density_fx, radius, inv_mass, dims):
    rho = density_fx.astype(np.float64) / 262144.0
    inv_rho0 = np.power(2.0 * radius.astype(np.float64), dims) * inv_mass
    e = np.abs(rho * inv_rho0 - 1.0)
    return float(e.mean()), float(e.max())


def hundred_substeps(side=20, n_sub=100):
    sc = scenes.uniform_block(side, jitter=0.2, shuffle=True, wall_gap=3.0)
    s = orc.default_settings()
    cap = sc.n * 80
    st = orc.State(**{k: v.copy() for k, v in sc.arrays.items()})
    ctx = apbf_b200.Context(dims=3)
    sim = apbf_b200.Sim(ctx, sc, neighbor_capacity=cap, integrate=True)
    sim.upload(sc.arrays)
    rows = []
    out = apbf_b200.empty_host_arrays(sc.n)
    for step in range(n_sub):
        ep = orc.substep(st, s, dims=3, basic_pbf=True, solver_iterations=4, min_pos=sc.min_pos, max_pos=sc.max_pos,
                         res_log2=sc.res_log2, box_min4=sc.box_min, box_max4=sc.box_max, cap=cap, integrate=True)
        sim.substep(1)
        if step % 10 == 9 or step == 0:
            # density of the end-of-substep state, evaluated by the oracle for both arms (same measuring stick)
            sim.download(out)
            tmp = st.copy()
            a = orc.incompressibility_apply(tmp, s, 3, ep, want_aux=True)
            g_state = orc.State(**{k: v.copy() for k, v in out.items()})
            gp = orc.green_apply(g_state, s, 3, 1.0, sc.min_pos, sc.max_pos, sc.res_log2, cap)
            b = orc.incompressibility_apply(g_state.copy(), s, 3, gp, want_aux=True)
            em, ex = density_stats(a["density"], st.radius, st.inverse_mass, 3)
            gm, gx = density_stats(b["density"], g_state.radius, g_state.inverse_mass, 3)
            dpos = np.abs(np.sort(out["position"][:, :3].astype(np.int64), axis=0) - np.sort(st.position[:, :3].astype(np.int64), axis=0))
            rows.append({"substep": step + 1, "oracle_mean": em, "oracle_max": ex, "cuda_mean": gm, "cuda_max": gx,
                         "pairs_oracle": int(len(ep)), "pairs_cuda": int(sim.neighbor_count()),
                         "sorted_coord_diff_p99_units": float(np.percentile(dpos, 99))})
    return {"scene": sc.name, "n": sc.n, "substeps": n_sub, "rows": rows}


def main():
    rep = {"one_iteration": [], "hundred_substeps": None}
    rep["one_iteration"].append(one_iteration(scenes.uniform_block(24, jitter=0.25, shuffle=True)))
    rep["one_iteration"].append(one_iteration(scenes.uniform_block(24, jitter=0.25, shuffle=True), hk=0, gk=0))
    rep["one_iteration"].append(one_iteration(scenes.uniform_block(24, jitter=0.25, shuffle=True), hk=2, gk=2))
    rep["one_iteration"].append(one_iteration(scenes.waterdrop(16, jitter=0.2), cap_per=300))
    rep["hundred_substeps"] = hundred_substeps()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_report.json"), "w") as f:
        json.dump(rep, f, indent=1)
    print(json.dumps(rep, indent=1))


if __name__ == "__main__":
    main()
