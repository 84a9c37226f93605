"""The tile formulation of the quadcopter KKT sweep (obca_quad.cuh, kkt_solve_block) restated in numpy, with the register fragments of
mma.sync.m8n8k4.f64 emulated lane by lane, against the plain dense stage step of the same elimination (kkt_dense):

    F = [Phi | r~] (20 x 24), Q with the stage gradient as column 21, T = P F (+ p on column 21), H = Q + F' T on the six upper tiles,
    gains from the LDL' of the 4 x 4 pivot block (feed-forward = column 21), P' = Hss + Hsu K mirrored, p' = column 21.

This is the index arithmetic of the kernel (leading dimensions, tile lists, which fragment element lives in which lane) checked on the
CPU; the kernel itself is checked on the GPU against the host recursion (tests/test_quadcopter.py)."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NS, NU, NY, IW, IT, IU = 17, 4, 21, 12, 16, 17
LDP, LDF, LDT, LDQ, LDH, LDK = 20, 24, 24, 24, 28, 24


def jac_pattern():
    src = open(os.path.join(ROOT, "obca_b200", "csrc", "obca_quad_dyn_gen.cuh")).read()
    tab = lambda name: [int(v) for v in re.search(r"#define %s \{([^}]*)\}" % name, src).group(1).split(",")]
    return tab("OBCA_QD_J_ROW"), tab("OBCA_QD_J_COL")


def dmma(Afr, Bfr, Cfr):
    """D = A (8x4) B (4x8) + C (8x8) with the fragment layout of mma.m8n8k4.f64: lane l holds A[l >> 2][l & 3], B[l & 3][l >> 2] and
    C / D [l >> 2][2 (l & 3) + {0, 1}]."""
    A = np.zeros((8, 4)); B = np.zeros((4, 8)); Cm = np.zeros((8, 8))
    for l in range(32):
        g, t = l >> 2, l & 3
        A[g, t] = Afr[l]; B[t, g] = Bfr[l]; Cm[g, 2 * t] = Cfr[l][0]; Cm[g, 2 * t + 1] = Cfr[l][1]
    D = A @ B + Cm
    return [[D[l >> 2, 2 * (l & 3)], D[l >> 2, 2 * (l & 3) + 1]] for l in range(32)]


def test_tile_formulation_equals_dense_stage_step():
    rng = np.random.default_rng(1)
    jr, jc = jac_pattern()
    assert len(jr) == 66 and max(jr) == 11 and max(jc) == 20
    for trial in range(3):
        # ---- stage data: value function of stage k+1, stage model, dynamics Jacobian and residual ----
        G = rng.normal(size=(NS, NS)); P = G @ G.T + NS * np.eye(NS); p = rng.normal(size=NS)
        G = rng.normal(size=(NY, NY)); Q = G @ G.T + NY * np.eye(NY); q = rng.normal(size=NY)
        Jv = rng.normal(size=66) * 0.3; r12 = rng.normal(size=12)
        Phi = np.zeros((NS, NY))
        for e in range(66):
            Phi[jr[e], jc[e]] += Jv[e]
        for j in range(NU):
            Phi[IW + j, IU + j] = 1.0                                   # w+ = u
        Phi[IT, IT] = 1.0                                               # t+ = t
        rt = np.zeros(NS); rt[:12] = r12
        # ---- dense reference (kkt_dense): g = p + P r~, H = Q + Phi' P Phi, hv = q + Phi' g, gains, value-function update ----
        g = p + P @ rt
        H = Q + Phi.T @ P @ Phi; hv = q + Phi.T @ g
        Huu = H[IU:, IU:]; K = -np.linalg.solve(Huu, H[IU:, :NS]); kf = -np.linalg.solve(Huu, hv[IU:])
        Pn = H[:NS, :NS] + H[:NS, IU:] @ K; pn = hv[:NS] + H[:NS, IU:] @ kf
        # ---- shared-memory tiles of the kernel ----
        Pm = np.zeros((24, LDP)); Pm[:NS, :NS] = P
        pv = np.zeros(24); pv[:NS] = p
        Fb = np.zeros((20, LDF)); Fb[:NS, :NY] = Phi; Fb[:12, 21] = r12
        Qb = np.zeros((24, LDQ))
        for i in range(NY):
            for j in range(i, NY):
                Qb[i, j] = Q[i, j]                                      # upper triangle only, as the prefetch scatters it
            Qb[i, 21] = q[i]
        Tm = np.zeros((20, LDT)); Hm = np.full((24, LDH), np.nan); Km = np.zeros((4, LDK))
        # T = P F, nine tiles, five k-steps each
        for idx in range(9):
            m, n = idx // 3, idx % 3
            c = [[0.0, 0.0] for _ in range(32)]
            for ks in range(5):
                A = [Pm[m * 8 + (l >> 2), ks * 4 + (l & 3)] for l in range(32)]
                B = [Fb[ks * 4 + (l & 3), n * 8 + (l >> 2)] for l in range(32)]
                c = dmma(A, B, c)
            for l in range(32):
                gid, tig = l >> 2, l & 3
                row = m * 8 + gid
                if row < 20:
                    c0, c1 = c[l]
                    if n == 2 and tig == 2:
                        c1 += pv[row]
                    Tm[row, n * 8 + tig * 2] = c0; Tm[row, n * 8 + tig * 2 + 1] = c1
        assert np.abs(Tm[:NS, :NY] - P @ Phi).max() < 1e-11 and np.abs(Tm[:NS, 21] - g).max() < 1e-11
        # H = Q + F' T on the six upper tiles
        tiles = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]
        for m, n in tiles:
            c = [[Qb[m * 8 + (l >> 2), n * 8 + (l & 3) * 2], Qb[m * 8 + (l >> 2), n * 8 + (l & 3) * 2 + 1]] for l in range(32)]
            for ks in range(5):
                A = [Fb[ks * 4 + (l & 3), m * 8 + (l >> 2)] for l in range(32)]
                B = [Tm[ks * 4 + (l & 3), n * 8 + (l >> 2)] for l in range(32)]
                c = dmma(A, B, c)
            for l in range(32):
                Hm[m * 8 + (l >> 2), n * 8 + (l & 3) * 2] = c[l][0]; Hm[m * 8 + (l >> 2), n * 8 + (l & 3) * 2 + 1] = c[l][1]
        iu = np.triu_indices(NY)
        assert np.abs(Hm[:NY, :NY][iu] - H[iu]).max() < 1e-10           # the upper triangle is what the later phases read
        assert np.abs(Hm[:NY, 21] - hv).max() < 1e-10
        # gains: one column per thread, LDL' of the pivot block read from the UPPER triangle, feed-forward stored as column 21
        for c_ in range(NS + 1):
            L = np.zeros((4, 4)); dd = np.zeros(4); di = np.zeros(4)
            for a in range(4):
                for b in range(a):
                    acc = Hm[IU + b, IU + a]
                    for l in range(b):
                        acc -= L[a, l] * L[b, l] * dd[l]
                    L[a, b] = acc * di[b]
                d = Hm[IU + a, IU + a]
                for l in range(a):
                    d -= L[a, l] ** 2 * dd[l]
                assert d > 0
                dd[a] = d; di[a] = 1.0 / d
            y4 = np.zeros(4); k4 = np.zeros(4)
            for a in range(4):
                acc = Hm[c_, IU + a] if c_ < NS else Hm[IU + a, 21]
                for l in range(a):
                    acc -= L[a, l] * y4[l]
                y4[a] = acc
            for a in range(3, -1, -1):
                acc = y4[a] * di[a]
                for l in range(a + 1, 4):
                    acc -= L[l, a] * k4[l]
                k4[a] = acc
            Km[:, c_ if c_ < NS else 21] = -k4
        assert np.abs(Km[:, :NS] - K).max() < 1e-10 and np.abs(Km[:, 21] - kf).max() < 1e-10
        # P' = Hss + Hsu K on the upper tiles, mirrored; column 21 is p'
        Pn_t = np.zeros((24, LDP)); pn_t = np.zeros(24)
        for m, n in tiles:
            c = [[Hm[m * 8 + (l >> 2), n * 8 + (l & 3) * 2], Hm[m * 8 + (l >> 2), n * 8 + (l & 3) * 2 + 1]] for l in range(32)]
            A = [Hm[m * 8 + (l >> 2), IU + (l & 3)] for l in range(32)]
            B = [Km[l & 3, n * 8 + (l >> 2)] for l in range(32)]
            with np.errstate(invalid="ignore"):
                c = dmma(np.nan_to_num(A), B, np.nan_to_num(np.array(c)).tolist())
            for l in range(32):
                a, b = m * 8 + (l >> 2), n * 8 + (l & 3) * 2
                c0, c1 = c[l]
                if a < NS:
                    if b < NS and a <= b:
                        Pn_t[a, b] = c0; Pn_t[b, a] = c0
                    if b + 1 < NS and a <= b + 1:
                        Pn_t[a, b + 1] = c1; Pn_t[b + 1, a] = c1
                    if b + 1 == 21:
                        pn_t[a] = c1
        assert np.abs(Pn_t[:NS, :NS] - Pn).max() < 1e-9 and np.abs(pn_t[:NS] - pn).max() < 1e-9
        assert np.abs(Pn_t[:NS, :NS] - Pn_t[:NS, :NS].T).max() == 0.0    # mirrored: exactly symmetric
