"""numpy MODEL of the device QP algorithm (test infrastructure, CPU only).

This mirrors, operation for operation, what the CUDA fast-path kernel does for one agent:
structured Mehrotra primal-dual interior point on

    min  sum_x [ 1/2 a_x' T a_x + q (lamK'a_x - e_x)^2 - 2 s ao_x a_x[0] ]  + sum_j (eps_j^2 + term eps_j)
    s.t. |a| <= alim,   d_j . y - dist_j eps_j >= r_j,   lb <= eps_j <= 0,     y_x = lamC' a_x

where T = 2(s Delta'Delta + I) is tridiagonal (K x K, per axis), lamK / lamC are rows of the
accel->position map, followed by an exact active-set "polish" with a KKT check.
It exists so that the maths can be validated against the oracle on CPU (`-m "not gpu"`), and as
executable documentation of the kernel.  It is never imported by the product.
"""
from __future__ import annotations

import numpy as np


def lam_rows(h: float, K: int):
    """lam[k, j] = A_p(3k+d, 3j+d) (getPosMat.m) via the reference recurrence; tt[k] = (k+1)h summed."""
    lam = np.zeros((K, K))
    pc = np.zeros(K)
    vc = np.zeros(K)
    tt = np.zeros(K)
    t = 0.0
    for k in range(K):
        pc = pc + h * vc
        pc[k] += h * h / 2
        vc[k] += h
        t += h
        lam[k] = pc
        tt[k] = t
    return lam, tt


def tridiag_T(s: float, K: int):
    """T = 2 (s Delta'Delta + I): diag, offdiag."""
    diag = np.full(K, 2.0 * (2.0 * s + 1.0))
    diag[K - 1] = 2.0 * (s + 1.0)
    off = np.full(K - 1, -2.0 * s)
    return diag, off


def pcr_solve(diag, off, rhs):
    """Parallel cyclic reduction for a symmetric tridiagonal system, 'lane = row' formulation.
    diag (K,), off (K-1,) with off[i] coupling rows i,i+1; rhs (K, nrhs).  Mirrors the kernel:
    log2 steps, each lane touches lanes i-d, i+d only."""
    K = len(diag)
    n = 1
    while n < K:
        n *= 2
    b = np.ones(n)
    b[:K] = diag
    a = np.zeros(n)          # coupling to row i-stride
    a[1:K] = off
    d = np.zeros((n, rhs.shape[1]))
    d[:K] = rhs
    stride = 1
    while stride < n:
        inv = 1.0 / b
        up = lambda v, fill: np.concatenate([np.full((stride,) + v.shape[1:], fill), v[:-stride]])
        dn = lambda v, fill: np.concatenate([v[stride:], np.full((stride,) + v.shape[1:], fill)])
        inv_m, inv_p = up(inv, 1.0), dn(inv, 1.0)
        a_m, a_p = up(a, 0.0), dn(a, 0.0)      # a_p = c_i (coupling to row i+stride)
        d_m, d_p = up(d, 0.0), dn(d, 0.0)
        alpha = -a * inv_m
        gamma = -a_p * inv_p
        b = b + alpha * a + gamma * a_p
        d = d + alpha[:, None] * d_m + gamma[:, None] * d_p
        a = alpha * a_m
        stride *= 2
    return (d / b[:, None])[:K]


class StructuredQP:
    """One agent's QP in structured form."""

    def __init__(self, K, h, q, s, alim, e, ao, kctr, d, dist, r, term, slack_lb):
        self.K, self.h, self.q, self.s, self.alim = K, h, q, s, alim
        self.e = np.asarray(e, float)          # (3,) pf - (po + t_K vo)
        self.ao = np.asarray(ao, float)        # (3,)
        self.lam, self.tt = lam_rows(h, K)
        self.lamK = self.lam[K - 1]
        self.nv = 0 if d is None else len(dist)
        self.d = np.zeros((0, 3)) if d is None else np.asarray(d, float).reshape(-1, 3)
        self.dist = np.zeros(0) if d is None else np.asarray(dist, float)
        self.r = np.zeros(0) if d is None else np.asarray(r, float)
        self.lamC = self.lam[kctr - 1] if self.nv else np.zeros(K)
        self.term, self.lb = term, slack_lb
        self.Td, self.To = tridiag_T(s, K)
        self.add_one_row = True

    # -- gradient of the smooth a-cost, per axis; a is (K,3)
    def grad_a(self, a):
        Ta = self.Td[:, None] * a
        Ta[1:] += self.To[:, None] * a[:-1]
        Ta[:-1] += self.To[:, None] * a[1:]
        tau = self.lamK @ a
        g = Ta + 2 * self.q * np.outer(self.lamK, tau - self.e)
        g[0] -= 2 * self.s * self.ao
        return g

    def objective(self, a, eps):
        tau = self.lamK @ a
        da = np.diff(a, axis=0, prepend=self.ao[None, :])
        v = self.q * ((tau - self.e) ** 2).sum() + (a ** 2).sum() + self.s * (da ** 2).sum()
        return v + (eps ** 2).sum() + self.term * eps.sum()

    # -- the structured linear solve shared by PDIP and polish ------------------------------
    def solve(self, Dadd, fixed, fixval, rhs, M3, m3, hard_d=None, hard_rho=None):
        """Solve for a (K,3):
              (T + Dadd) a + 2q lamK (lamK'a) + lamC (M3 y + m3 - Dh' nu) = rhs   on free entries
              a = fixval on fixed entries,  Dh y = rho  (hard rows),  y = lamC' a.
        rhs must NOT contain the 2q lamK e term's dependence on a (it is a plain vector).
        Returns a, y, nu or None if the small system is singular."""
        K, q = self.K, self.q
        a = np.zeros((K, 3))
        kKK, kKc, kcc, rK, rc = (np.zeros(3) for _ in range(5))
        W = []
        for x in range(3):
            fx = fixed[:, x]
            diag = np.where(fx, 1.0, self.Td + Dadd[:, x])
            off = np.where(fx[:-1] | fx[1:], 0.0, self.To)
            rr = np.where(fx, fixval[:, x], rhs[:, x])
            # move couplings to fixed neighbours to the right-hand side (keeps symmetry)
            fv = np.where(fx, fixval[:, x], 0.0)
            rr[1:] -= np.where(~fx[1:], self.To * fv[:-1], 0.0)
            rr[:-1] -= np.where(~fx[:-1], self.To * fv[1:], 0.0)
            lK = np.where(fx, 0.0, self.lamK)
            lC = np.where(fx, 0.0, self.lamC)
            w = pcr_solve(diag, off, np.stack([rr, lK, lC], axis=1))
            W.append(w)
            kKK[x] = self.lamK @ w[:, 1]
            kKc[x] = self.lamK @ w[:, 2]
            kcc[x] = self.lamC @ w[:, 2]
            rK[x] = self.lamK @ w[:, 0]
            rc[x] = self.lamC @ w[:, 0]
        den = 1.0 + 2 * q * kKK
        rho_hat = rc - 2 * q * kKc * rK / den
        kap_hat = kcc - 2 * q * kKc ** 2 / den
        nh = 0 if hard_d is None else len(hard_rho)
        # (I + Khat M3) y - Khat Dh' nu = rho_hat - Khat m3 ;  Dh y = rho
        S = np.zeros((3 + nh, 3 + nh))
        t = np.zeros(3 + nh)
        S[:3, :3] = np.eye(3) + kap_hat[:, None] * M3
        t[:3] = rho_hat - kap_hat * m3
        if nh:
            S[:3, 3:] = -kap_hat[:, None] * hard_d.T
            S[3:, :3] = hard_d
            t[3:] = hard_rho
        try:
            if nh and np.linalg.cond(S) > 1e12:
                return None
            sol = np.linalg.solve(S, t)
        except np.linalg.LinAlgError:
            return None
        y, nu = sol[:3], sol[3:]
        g = M3 @ y + m3 - (hard_d.T @ nu if nh else 0.0)
        tau = (rK - g * kKc) / den
        for x in range(3):
            w = W[x]
            a[:, x] = w[:, 0] - 2 * q * tau[x] * w[:, 1] - g[x] * w[:, 2]
        return a, y, nu

    # -- unconstrained optimum (closed form gain in the kernel) -------------------------------
    def unconstrained(self):
        K = self.K
        rhs = 2 * self.q * np.outer(self.lamK, self.e)
        rhs[0] += 2 * self.s * self.ao
        z = np.zeros((K, 3))
        a, _, _ = self.solve(z, z.astype(bool), z, rhs, np.zeros((3, 3)), np.zeros(3))
        return a

    # -- exact polish for a guessed active set -----------------------------------------------
    def polish(self, act_u, act_l, row_state, tol=1e-9):
        """act_u/act_l (K,3) bool: a fixed at +alim / -alim.
        row_state[j]: 0 inactive (eps=0), 1 penalty (eps free), 2 hard at eps=lb, 3 hard at eps=0.
        Returns (a, eps, ok, next_guess) where next_guess = (act_u, act_l, row_state) corrected by
        the primal-dual active-set rule (None if the small system was singular)."""
        K = self.K
        fixed = act_u | act_l
        fixval = np.where(act_u, self.alim, -self.alim)
        rhs = 2 * self.q * np.outer(self.lamK, self.e)
        rhs[0] += 2 * self.s * self.ao
        pen = row_state == 1
        hard = row_state >= 2
        if hard.sum() > 3:
            return None, None, False, None
        d, dist, r = self.d, self.dist, self.r
        M3 = (2.0 / dist[pen] ** 2 * d[pen].T) @ d[pen] if pen.any() else np.zeros((3, 3))
        m3 = ((self.term / dist[pen] - 2 * r[pen] / dist[pen] ** 2)[:, None] * d[pen]).sum(0) if pen.any() else np.zeros(3)
        hd = d[hard]
        hrho = np.where(row_state[hard] == 2, r[hard] + dist[hard] * self.lb, r[hard])
        res = self.solve(np.zeros((K, 3)), fixed, fixval, rhs, M3, m3, hd if hard.any() else None,
                         hrho if hard.any() else None)
        if res is None:
            return None, None, False, None
        a, y, nu = res
        # ---- KKT verification + corrected guess
        eps = np.zeros(self.nv)
        zr = np.zeros(self.nv)
        if self.nv:
            sr0 = d @ y - r                       # row slack with eps = 0
            eps[pen] = sr0[pen] / dist[pen]
            eps[row_state == 2] = self.lb
            zr[pen] = -(self.term + 2 * eps[pen]) / dist[pen]
            zr[hard] = nu
        g = self.grad_a(a)
        if self.nv:
            g -= np.outer(self.lamC, zr @ d)
        gt = tol * (1 + np.abs(g).max())
        n_u = (act_u & ~(g > gt)) | (~fixed & (a > self.alim + tol))
        n_l = (act_l & ~(g < -gt)) | (~fixed & (a < -self.alim - tol))
        ok = np.array_equal(n_u, act_u) and np.array_equal(n_l, act_l)
        rs = row_state.copy()
        if self.nv:
            zt = tol * (1 + abs(self.term))
            s0, s1, s2, s3 = (row_state == i for i in range(4))
            viol = s0 & (sr0 < -tol)
            if viol.any():
                if self.add_one_row:
                    jworst = np.argmin(np.where(viol, sr0 / np.linalg.norm(d, axis=1), np.inf))
                    rs[jworst] = 3
                else:
                    rs[viol] = 3
            rs[s1 & (eps > tol)] = 3
            rs[s1 & (eps < self.lb - tol)] = 2
            zel = 2 * self.lb + self.term + dist * zr
            zeu = -(self.term + dist * zr)
            rs[s2 & (zel < -zt)] = 1
            rs[s2 & (zr < -zt)] = 0
            rs[s3 & (zeu < -zt)] = 1
            rs[s3 & (zr < -zt)] = 0
            ok = ok and np.array_equal(rs, row_state)
        return a, eps, ok, (n_u, n_l, rs)

    def pdas(self, guess, max_iter=4):
        """primal-dual active-set iterations from a guess; returns (a, eps, ok, iters)."""
        for it in range(max_iter):
            a, eps, ok, guess = self.polish(*guess)
            if ok:
                return a, eps, True, it + 1
            if guess is None:
                return None, None, False, it + 1
        return None, None, False, max_iter

    # -- Mehrotra PDIP with early polish ------------------------------------------------------
    def pdip(self, max_iter=60, tol=1e-10, polish_mu=1e-3, verbose=False, mu0_rel=1e-2, sr_floor=1e-2, pdas_iters=3, cold_pdas=4):
        K, nv, alim = self.K, self.nv, self.alim
        d, dist, r, lb, term = self.d, self.dist, self.r, self.lb, self.term
        info = dict(iters=0, polished=False, polish_tries=0, status="maxiter")
        # stage 1/2: unconstrained optimum, then cold-start active-set iterations
        a_unc = self.unconstrained()
        if cold_pdas:
            zb = np.zeros((K, 3), bool)
            ap, ep, ok, npol = self.pdas((zb, zb, np.zeros(nv, int)), cold_pdas)
            info["cold_solves"] = npol
            if ok:
                info.update(polished=True, status="ok", cold=True)
                return ap, ep, info
        a = np.clip(a_unc, -0.9 * alim, 0.9 * alim)
        fscale = max(1.0, np.abs(self.grad_a(np.zeros((K, 3)))).max())
        mu0 = mu0_rel * fscale
        zu = mu0 / (alim - a)
        zl = mu0 / (alim + a)
        y = self.lamC @ a
        if nv:
            # eps just below its upper bound 0, its dual balancing the linear penalty `term`
            eps = np.full(nv, max(-mu0 / abs(term), 0.5 * lb))
            zeu = np.full(nv, abs(term))
            zel = mu0 / (eps - lb)
            sr = np.maximum(d @ y - dist * eps - r, sr_floor)
            zr = mu0 / sr
        else:
            eps = sr = zr = zeu = zel = np.zeros(0)
        m_tot = 6 * K + 3 * nv
        for it in range(max_iter):
            su, sl = alim - a, alim + a
            seu, sel = -eps, eps - lb
            y = self.lamC @ a
            ra = self.grad_a(a) + zu - zl
            if nv:
                ra -= np.outer(self.lamC, zr @ d)
                re = 2 * eps + term + zeu - zel + dist * zr
                rpr = -(d @ y) + dist * eps + sr + r
            else:
                re = rpr = np.zeros(0)
            mu = ((su * zu).sum() + (sl * zl).sum() + (seu * zeu).sum() + (sel * zel).sum() + (sr * zr).sum()) / m_tot
            res_d = max(np.abs(ra).max(), np.abs(re).max() if nv else 0.0) / fscale
            res_p = np.abs(rpr).max() if nv else 0.0
            info.update(iters=it, mu=mu, res_d=res_d, res_p=res_p)
            if verbose:
                print(f"it {it:2d} mu {mu:.3e} rd {res_d:.3e} rp {res_p:.3e}")
            # --- early polish attempt
            if mu < polish_mu * fscale and res_p < 1e-3:
                info["polish_tries"] += 1
                act_u, act_l = zu > su, zl > sl
                rs = np.zeros(nv, int)
                if nv:
                    active = zr > sr
                    at0 = zeu > seu
                    atlb = zel > sel
                    rs[active & ~at0 & ~atlb] = 1
                    rs[active & atlb] = 2
                    rs[active & at0 & ~atlb] = 3
                ap, ep, ok, npol = self.pdas((act_u, act_l, rs), pdas_iters)
                info["polish_solves"] = info.get("polish_solves", 0) + npol
                if ok:
                    info.update(polished=True, status="ok")
                    return ap, ep, info
            if max(res_d, res_p) < tol and mu < tol * fscale:
                info["status"] = "ok"
                return a, eps, info
            # infeasibility certificate (rows + box): z_r >= 0 with min over box of (sum z_j d_j).y' ...
            if nv and it > 5:
                zc = zr / zr.sum() if zr.sum() > 0 else zr
                cvec = np.outer(self.lamC, zc @ d)              # coefficient of a in  sum z_j d_j.y
                best = np.abs(cvec).sum() * alim                # max over the box
                need = zc @ (r + dist * lb)                     # rows relaxed to eps = lb
                if best < need - 1e-9 * (1 + abs(need)):
                    info["status"] = "infeasible"
                    return None, None, info
            # --- Newton system
            wu, wl = zu / su, zl / sl
            Dadd = wu + wl
            if nv:
                w_u, w_l, w_r = zeu / seu, zel / sel, zr / sr
                Theta = 2 + w_u + w_l + dist ** 2 * w_r
                wt = w_r * (2 + w_u + w_l) / Theta
                M3 = (wt * d.T) @ d
            else:
                M3 = np.zeros((3, 3))
            nofix = np.zeros((K, 3), bool)

            def newton(rcu, rcl, rceu, rcel, rcr):
                rhs = -ra + rcu / su - rcl / sl
                if nv:
                    g0 = -re + rceu / seu - rcel / sel + dist * (rcr / sr - w_r * rpr)
                    c0 = -rcr / sr + w_r * rpr
                    zeta = c0 + w_r * dist * g0 / Theta
                    m3 = -(zeta @ d)        # enters as + lamC*(M3 y + m3) on the lhs
                else:
                    m3 = np.zeros(3)
                da, dy, _ = self.solve(Dadd, nofix, np.zeros((K, 3)), rhs, M3, m3)
                dzu = -rcu / su + wu * da
                dzl = -rcl / sl - wl * da
                if nv:
                    dyv = d @ dy
                    deps = (g0 + dist * w_r * dyv) / Theta
                    dzr = zeta - wt * dyv
                    dsr = -rpr + dyv - dist * deps
                    dzeu = -rceu / seu + w_u * deps
                    dzel = -rcel / sel - w_l * deps
                else:
                    deps = dzr = dsr = dzeu = dzel = np.zeros(0)
                return da, deps, dsr, dzu, dzl, dzr, dzeu, dzel

            def maxstep(v, dv):
                neg = dv < 0
                return min(1.0, (-v[neg] / dv[neg]).min()) if neg.any() else 1.0

            def steps(da, deps, dsr, dzu, dzl, dzr, dzeu, dzel):
                ap = min(maxstep(su, -da), maxstep(sl, da), maxstep(seu, -deps), maxstep(sel, deps), maxstep(sr, dsr))
                ad = min(maxstep(zu, dzu), maxstep(zl, dzl), maxstep(zr, dzr), maxstep(zeu, dzeu), maxstep(zel, dzel))
                return min(ap, ad)

            aff = newton(su * zu, sl * zl, seu * zeu, sel * zel, sr * zr)
            al = steps(*aff)
            da, deps, dsr, dzu, dzl, dzr, dzeu, dzel = aff
            mua = (((su - al * da) * (zu + al * dzu)).sum() + ((sl + al * da) * (zl + al * dzl)).sum()
                   + ((seu - al * deps) * (zeu + al * dzeu)).sum() + ((sel + al * deps) * (zel + al * dzel)).sum()
                   + ((sr + al * dsr) * (zr + al * dzr)).sum()) / m_tot
            sigma = (mua / mu) ** 3
            sm = sigma * mu
            cor = newton(su * zu + (-da) * dzu - sm, sl * zl + da * dzl - sm,
                         seu * zeu + (-deps) * dzeu - sm, sel * zel + deps * dzel - sm,
                         sr * zr + dsr * dzr - sm)
            al = min(1.0, 0.99 * steps(*cor)) if steps(*cor) < 1.0 else 1.0
            al = min(al, 0.999999) if al == 1.0 else al
            da, deps, dsr, dzu, dzl, dzr, dzeu, dzel = cor
            a = a + al * da
            zu, zl = zu + al * dzu, zl + al * dzl
            if nv:
                eps = eps + al * deps
                sr = sr + al * dsr
                zr, zeu, zel = zr + al * dzr, zeu + al * dzeu, zel + al * dzel
        return a, eps, info
