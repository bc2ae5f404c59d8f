#!/usr/bin/env python
"""CPU emulation of the 4-lanes-per-window Jacobi used by the fused kernel's eigensolver warp
(gr-baz_b200/csrc/music_fused.cuh::herm_eig4_coop) against the sequential one-lane solver
(music_kernels.cuh::herm_eig_body<4, true>): same slot layout (slot t of lane j = row j ^ t of column j), same
shuffle pattern, same order of operations - the two must agree bit for bit (no FMA on either side here)."""
import numpy as np


def params(gr, gi, app, aqq):
    gg = gr * gr + gi * gi
    nz = gg > 0.0
    rg = 1.0 / np.sqrt(gg) if nz else 0.0
    g = gg * rg
    er, ei = gr * rg, gi * rg
    theta = 0.5 * (aqq - app) * rg
    theta = min(max(theta, -1e150), 1e150)
    q1 = theta * theta + 1.0
    sq = q1 * (1.0 / np.sqrt(q1))
    t = 1.0 / (abs(theta) + sq)
    t = float(np.copysign(t, theta)) if nz else 0.0
    c = 1.0 / np.sqrt(t * t + 1.0)
    s = t * c
    return c, s * er, s * ei, t * g


def mix(c, swr, swi, kpr, kpi, kqr, kqi):
    npr = c * kpr - (swr * kqr + swi * kqi)
    npi = c * kpi - (swr * kqi - swi * kqr)
    nqr = c * kqr + (swr * kpr - swi * kpi)
    nqi = c * kqi + (swr * kpi + swi * kpr)
    return npr, npi, nqr, nqi


def rotate_seq(Ar, Ai, Vr, Vi, p, q):
    c, swr, swi, tg = params(Ar[p][q], Ai[p][q], Ar[p][p], Ar[q][q])
    app, aqq = Ar[p][p], Ar[q][q]
    for k in range(4):
        if k in (p, q):
            continue
        npr, npi, nqr, nqi = mix(c, swr, swi, Ar[k][p], Ai[k][p], Ar[k][q], Ai[k][q])
        Ar[k][p], Ai[k][p], Ar[k][q], Ai[k][q] = npr, npi, nqr, nqi
        Ar[p][k], Ai[p][k], Ar[q][k], Ai[q][k] = npr, -npi, nqr, -nqi
    Ar[p][p], Ai[p][p] = app - tg, 0.0
    Ar[q][q], Ai[q][q] = aqq + tg, 0.0
    Ar[p][q] = Ai[p][q] = Ar[q][p] = Ai[q][p] = 0.0
    for k in range(4):
        npr, npi, nqr, nqi = mix(c, swr, swi, Vr[k][p], Vi[k][p], Vr[k][q], Vi[k][q])
        Vr[k][p], Vi[k][p], Vr[k][q], Vi[k][q] = npr, npi, nqr, nqi


def eig_seq(R):
    Ar, Ai = R.real.copy(), R.imag.copy()
    Vr, Vi = np.eye(4), np.zeros((4, 4))
    sweeps = 0
    for _ in range(60):
        fro = off = 0.0
        for i in range(4):
            for j in range(4):
                e2 = Ar[i][j] * Ar[i][j] + Ai[i][j] * Ai[i][j]
                fro += e2
                if i != j:
                    off += e2
        if off <= 1e-32 * fro or off == 0.0:
            break
        for p, q in ((0, 1), (2, 3), (0, 2), (1, 3), (0, 3), (1, 2)):
            rotate_seq(Ar, Ai, Vr, Vi, p, q)
        sweeps += 1
    return Ar, Ai, Vr, Vi, sweeps


def eig_coop(R):
    """4 lanes; lane j holds column j of A and V in XOR-relative slots: slot t <-> row j ^ t."""
    ar = [[R[j ^ t, j].real for t in range(4)] for j in range(4)]
    ai = [[R[j ^ t, j].imag for t in range(4)] for j in range(4)]
    vr = [[1.0 if t == 0 else 0.0 for t in range(4)] for j in range(4)]
    vi = [[0.0] * 4 for j in range(4)]
    sweeps = 0
    for _ in range(60):
        fro = sum(ar[j][t] ** 2 + ai[j][t] ** 2 for j in range(4) for t in range(4))
        off = sum(ar[j][t] ** 2 + ai[j][t] ** 2 for j in range(4) for t in range(1, 4))
        if off <= 1e-32 * fro or off == 0.0:
            break
        for x in (1, 2, 3):
            o = 2 if x == 1 else 1          # slot of one row of the other pair; the other one is o ^ x
            hb = 1 if x == 1 else 2         # highest bit of x: decides who is q in a pair
            # --- every lane: parameters of its own pair (needs the mate's diagonal = mate's slot 0)
            P = []
            for j in range(4):
                mate = j ^ x
                isq = (j & hb) != 0
                dm = ar[mate][0]            # shuffle
                if not isq:
                    app, aqq, gr, gi = ar[j][0], dm, ar[j][x], -ai[j][x]
                else:
                    app, aqq, gr, gi = dm, ar[j][0], ar[j][x], ai[j][x]
                P.append(params(gr, gi, app, aqq) + (app, aqq, isq))
            PO = [P[j ^ o] for j in range(4)]  # shuffle: parameters of the other pair

            def local(j, prm):
                """rotation of the OTHER pair seen from column j: rows j ^ o and j ^ o ^ x, the smaller row is 'p'"""
                c, swr, swi = prm[0], prm[1], prm[2]
                sp, sq_ = (o, o ^ x) if ((j ^ o) & hb) == 0 else (o ^ x, o)
                npr, npi, nqr, nqi = mix(c, swr, swi, ar[j][sp], -ai[j][sp], ar[j][sq_], -ai[j][sq_])
                ar[j][sp], ai[j][sp], ar[j][sq_], ai[j][sq_] = npr, -npi, nqr, -nqi

            first = [(j == 0) or (j == x) for j in range(4)]
            # phase 1: lanes of the second pair apply the first rotation to their column
            for j in range(4):
                if not first[j]:
                    local(j, PO[j])
            # phase 2: exchange the off-block rows with the mate (mate's slot for my row j ^ t is x ^ t ... same t)
            snap_r = [row[:] for row in ar]
            snap_i = [row[:] for row in ai]
            # phase 3: column mix with the own pair's parameters
            for j in range(4):
                mate = j ^ x
                c, swr, swi, tg, app, aqq, isq = P[j]
                for t in (o, o ^ x):
                    # my row j ^ t sits in the mate's slot (j ^ t) ^ mate = t ^ x
                    mr, mi = snap_r[mate][t ^ x], snap_i[mate][t ^ x]
                    if not isq:
                        npr, npi, _, _ = mix(c, swr, swi, snap_r[j][t], snap_i[j][t], mr, mi)
                        ar[j][t], ai[j][t] = npr, npi
                    else:
                        _, _, nqr, nqi = mix(c, swr, swi, mr, mi, snap_r[j][t], snap_i[j][t])
                        ar[j][t], ai[j][t] = nqr, nqi
                # own diagonal block
                if not isq:
                    ar[j][0], ai[j][0] = app - tg, 0.0
                else:
                    ar[j][0], ai[j][0] = aqq + tg, 0.0
                ar[j][x], ai[j][x] = 0.0, 0.0
            # phase 4: lanes of the first pair apply the second rotation to their column
            for j in range(4):
                if first[j]:
                    local(j, PO[j])
            # V: whole columns mix with the mate's column (same row k = j ^ t -> mate's slot t ^ x)
            svr = [row[:] for row in vr]
            svi = [row[:] for row in vi]
            for j in range(4):
                mate = j ^ x
                c, swr, swi, tg, app, aqq, isq = P[j]
                for t in range(4):
                    mr, mi = svr[mate][t ^ x], svi[mate][t ^ x]
                    if not isq:
                        vr[j][t], vi[j][t] = mix(c, swr, swi, svr[j][t], svi[j][t], mr, mi)[:2]
                    else:
                        vr[j][t], vi[j][t] = mix(c, swr, swi, mr, mi, svr[j][t], svi[j][t])[2:]
        sweeps += 1
    Ar = np.array([[ar[j][i ^ j] for j in range(4)] for i in range(4)])
    Ai = np.array([[ai[j][i ^ j] for j in range(4)] for i in range(4)])
    Vr = np.array([[vr[j][i ^ j] for j in range(4)] for i in range(4)])
    Vi = np.array([[vi[j][i ^ j] for j in range(4)] for i in range(4)])
    return Ar, Ai, Vr, Vi, sweeps


if __name__ == "__main__":
    rng = np.random.default_rng(7)
    worst = 0
    for trial in range(300):
        X = rng.standard_normal((4, 64)) + 1j * rng.standard_normal((4, 64))
        if trial % 3 == 0:  # one strong source, like the MUSIC windows
            a = np.exp(1j * rng.uniform(0, 6.28, 4))[:, None]
            X = a * (rng.standard_normal((1, 64)) + 1j * rng.standard_normal((1, 64))) * 10 + X
        R = X @ X.conj().T / 64
        R = (R + R.conj().T) / 2
        R[np.diag_indices(4)] = R[np.diag_indices(4)].real
        s = eig_seq(R)
        c = eig_coop(R)
        same = all(np.array_equal(a, b) for a, b in zip(s[:4], c[:4])) and s[4] == c[4]
        if not same:
            print("trial", trial, "MISMATCH sweeps", s[4], c[4], "max |dA|", np.abs(s[0] - c[0]).max(), "max |dV|", np.abs(s[2] - c[2]).max())
            worst += 1
    print("mismatching trials:", worst, "of 300")
