"""A second, independent restatement (numpy, array-at-a-time) of the RAM drift routines, written
from the Fortran text and not from oracle/ram_oracle.cpp: DRIFTPARA, DRIFTR, DRIFTP, DRIFTE, DRIFTMU
(src/ModRamDrift.f90:36-473).  The reference cannot be built here, so the
C++ oracle is "parity unpinned"; two restatements of different shape (scalar loops in C++,
whole-array expressions here) that agree bit for bit at least rule out transcription slips in
either.  IEEE double arithmetic, same operation order as the Fortran, no fused multiply-add.

Arrays use the reference's shapes with 0-based numpy indices: F2[S,I,J,K,L], fields [I,J(,L)]
with the radial ghost row at index NR.
"""
import numpy as np

OME = 7.3E-5
Q, CS = 1.602E-19, 2.998E8
FRAC_CFL = 0.8


def limiter_flux(Fm, Fp, Fn, Fn1, c, chat, sgn, beta):
    """FBND for every interface of an array of lines.  Fm=F(m), Fp=F(m+1), Fn=F(n), Fn1=F(n-1) with
    n = m+1-sgn already gathered by the caller.  (:246-259)"""
    X = Fp - Fm
    FUP = 0.5 * (Fm + Fp - sgn * X)
    with np.errstate(divide="ignore", invalid="ignore"):
        R = (Fn - Fn1) / X
        LIM = np.maximum(np.minimum(beta * R, 1.), np.minimum(R, beta))
        CORR = -0.5 * (chat - sgn) * X
        lim = FUP + LIM * CORR
    use = (np.abs(X) > 1.E-27) & (R > 0)
    return np.where(use, lim, FUP)


def driftpara(g, S, DTs):
    """P1(I), P2(I,K), EDOT(I,K) (:64-85); QS = species charge."""
    s = S - 1
    QS = float(g.QS[s])
    RLZ = g.RLZ[:g.NR]
    P1 = DTs / g.DPHI / 2 / g.MDR / RLZ
    P2 = DTs * g.EKEV[None, :] * 1000 * (g.GREL[s][None, :] + 1) / g.GREL[s][None, :] / (RLZ[:, None] ** 2) / g.DPHI / QS
    EDOT = g.EBND[None, :] * DTs / RLZ[:, None] * (g.GRBND[s][None, :] + 1) / g.GRBND[s][None, :] / 2.
    return QS, P1, P2, EDOT


def driftp(g, inp, F2, S, DTs, beta):
    """DRIFTP for species S; returns (new F2 of the species, DtDriftP)."""
    NR, NT, NE, NPA = g.NR, g.NT, g.NE, g.NPA
    s = S - 1
    QS, P1, P2, EDOT = driftpara(g, S, DTs)
    F = F2[s].copy()                       # [I,J,K,L]
    FN, FH, B, VT, EIR = inp.FNIS, inp.FNHS, inp.BNES, inp.VT, inp.EIR
    RLZ = g.RLZ[:NR]
    I = np.arange(1, NR)                   # Fortran I = 2..NR
    Jv = np.arange(1, NT)                  # Fortran J = 2..NT
    J1 = np.where(Jv == NT - 1, 1, Jv + 1)  # J1 = J+1, 2 when J = NT
    ix = np.ix_
    # 3-D pieces [I,J,L]
    GPA1 = (FN[ix(I, Jv)] + FN[ix(I, J1)] + (FN[ix(I + 1, J1)] + FN[ix(I + 1, Jv)] - FN[ix(I - 1, Jv)] - FN[ix(I - 1, J1)])
            * RLZ[I][:, None, None] / 2. / g.MDR)
    GPA2 = (RLZ[I][:, None, None] / 4. / g.MDR * (FN[ix(I, Jv)] + FN[ix(I, J1)] - 2 * FH[ix(I, Jv)] - 2 * FH[ix(I, J1)])
            * (B[ix(I + 1, J1)] + B[ix(I + 1, Jv)] - B[ix(I - 1, Jv)] - B[ix(I - 1, J1)])[:, :, None]
            / (B[ix(I, Jv)] + B[ix(I, J1)])[:, :, None])
    A = ((VT[ix(I + 1, Jv)] + VT[ix(I + 1, J1)] - VT[ix(I - 1, Jv)] - VT[ix(I - 1, J1)]) * P1[I][:, None])      # [I,J]
    Cterm = (EIR[ix(I, J1)] + EIR[ix(I, Jv)]) / RLZ[I][:, None] * DTs / g.DPHI                                   # [I,J]
    sB = (B[ix(I, Jv)] + B[ix(I, J1)])                                                                           # [I,J]
    # CDriftP[I,J,K,L]
    Bt = P2[I][:, None, :, None] * (GPA1 + GPA2)[:, :, None, :] / (FH[ix(I, Jv)] + FH[ix(I, J1)])[:, :, None, :]
    CD = (A[:, :, None, None] - Bt - Cterm[:, :, None, None]) / sB[:, :, None, None] + OME * DTs / g.DPHI
    inside = (inp.outsideMGNP[ix(I, Jv)] == 0)
    ctemp = np.maximum(np.abs(CD), 1E-10)
    dts = np.where(inside[:, :, None, None], FRAC_CFL * DTs / ctemp, np.inf)
    Dt = min(100000.0, float(dts.min()))
    sgn = np.where(CD < 0, -1.0, 1.0)
    Fl = F[1:]                              # [I(2..NR), J(1..NT), K, L]
    Fm, Fp = Fl[:, Jv], Fl[:, J1]
    # n = J+1-sgn (wrapped: N > NT -> N-NT+1): sgn=+1 -> n=J, n-1=J-1 ; sgn=-1 -> n=J+2, n-1=J+1
    N2 = Jv + 2
    N2 = np.where(N2 > NT - 1, N2 - NT + 1, N2)          # 0-based: index > NT-1 wraps to index-NT+1
    Fn = np.where(sgn > 0, Fm, Fl[:, N2])
    Fn1 = np.where(sgn > 0, Fl[:, Jv - 1], Fp)
    FB = limiter_flux(Fm, Fp, Fn, Fn1, CD, CD, sgn, beta)            # FBND(J), J = 2..NT
    # FBND(1) = FBND(NT), CDriftP(.,1,.,.) = CDriftP(.,NT,.,.)
    FBm1 = np.concatenate([FB[:, -1:], FB[:, :-1]], axis=1)          # FBND(J-1)
    CDm1 = np.concatenate([CD[:, -1:], CD[:, :-1]], axis=1)
    new = Fm - CD * FB + CDm1 * FBm1
    new = np.where(new < 0, 1E-15, new)
    out = F.copy()
    out[1:, 1:] = new
    out[1:, 0] = out[1:, NT - 1]
    return out, Dt


def drifte(g, inp, F2, S, DTs, beta):
    """DRIFTE for species S; returns (new F2 of the species, DtDriftE)."""
    NR, NT, NE, NPA = g.NR, g.NT, g.NE, g.NPA
    s = S - 1
    QS, P1, P2, EDOT = driftpara(g, S, DTs)
    F = F2[s].copy()
    FN, FH, B, VT, EIR, EIP = inp.FNIS, inp.FNHS, inp.BNES, inp.VT, inp.EIR, inp.EIP
    RLZ = g.RLZ[:NR]
    I = np.arange(1, NR)
    J = np.arange(0, NT)
    J0 = np.where(J == 0, NT - 2, J - 1)
    J2 = np.where(J == NT - 1, 1, J + 1)
    ix = np.ix_
    R1 = RLZ[I][:, None]
    DRD1 = (EIP[ix(I, J)] * R1 - (VT[ix(I, J2)] - VT[ix(I, J0)]) / 2. / g.DPHI) / B[ix(I, J)]
    DPD1 = OME * R1 + ((VT[ix(I + 1, J)] - VT[ix(I - 1, J)]) / 2 / g.MDR - EIR[ix(I, J)]) / B[ix(I, J)]
    Bc = B[ix(I, J)][:, :, None]
    R3 = R1[:, :, None]
    FNc, FHc = FN[ix(I, J)], FH[ix(I, J)]
    GPA = (1. - FNc / 2. / FHc) / Bc
    GPR1 = GPA * (B[ix(I + 1, J)] - B[ix(I - 1, J)])[:, :, None] / 2. / g.MDR
    GPR2 = -FNc / FHc / R3
    GPR3 = -(FN[ix(I + 1, J)] - FN[ix(I - 1, J)]) / 2. / g.MDR / FHc
    GPP1 = GPA * (B[ix(I, J2)] - B[ix(I, J0)])[:, :, None] / 2. / g.DPHI
    GPP2 = -(FN[ix(I, J2)] - FN[ix(I, J0)]) / 2. / g.DPHI / FHc
    DRD2 = ((FN[ix(I, J2)] - FN[ix(I, J0)]) / 2. / g.DPHI
            + (FNc - 2 * FHc) * (B[ix(I, J2)] - B[ix(I, J0)])[:, :, None] / 4 / Bc / g.DPHI)
    DPD2 = (FNc + (FN[ix(I + 1, J)] - FN[ix(I - 1, J)]) * R3 / 2 / g.MDR
            + R3 * (FNc - 2 * FHc) / 4 / g.MDR * (B[ix(I + 1, J)] - B[ix(I - 1, J)])[:, :, None] / Bc)
    dBdt1 = inp.dBdt[ix(I, J)][:, :, None] * (1. - FNc / 2. / FHc) * R3 / Bc
    dIdt1 = -inp.dIdt[ix(I, J)] * R3 / FHc
    eK = g.EBND * 1e3 * (g.GRBND[s] + 1) / 2 / g.GRBND[s]                  # [K]
    # [I,J,K,L]
    EDT1 = eK[None, None, :, None] / FHc[:, :, None, :] / R3[:, :, None, :] / Bc[:, :, None, :] / QS
    DRDT = DRD1[:, :, None, None] + EDT1 * DRD2[:, :, None, :] * R3[:, :, None, :]
    DPDT = DPD1[:, :, None, None] - EDT1 * DPD2[:, :, None, :]
    CD = EDOT[I][:, None, :, None] * ((GPR1 + GPR2 + GPR3)[:, :, None, :] * DRDT + (GPP1 + GPP2)[:, :, None, :] * DPDT
                                      + dBdt1[:, :, None, :] + dIdt1[:, :, None, :])
    inside = (inp.outsideMGNP[ix(I, J)] == 0)
    ctemp = np.maximum(np.abs(CD), 1E-10)
    dts = np.where(inside[:, :, None, None], FRAC_CFL * DTs * g.DE[None, None, :, None] / ctemp, np.inf)
    Dt = min(10000.0, float(dts.min()))
    # line buffer F(0:NE+2): F(1), F(0) ghosts, F(NE+1) = F(NE+2) = 0
    Fl = F[1:]                                                            # [I,J,K,L]
    G = g.GREL[s]
    EZERO = g.EKEV[0] - g.WE[0]
    GRZ = 1. + EZERO * 1000. * Q / g.RMAS[s] / CS / CS
    f1 = Fl[:, :, 1] * G[0] / G[1] * np.sqrt((G[1] ** 2 - 1) / (G[0] ** 2 - 1))
    f0 = f1 * GRZ / G[0] * np.sqrt((G[0] ** 2 - 1) / (GRZ ** 2 - 1))
    z = np.zeros_like(f1)
    buf = np.concatenate([f0[:, :, None], f1[:, :, None], Fl[:, :, 1:], z[:, :, None], z[:, :, None]], axis=2)   # index = Fortran K
    K = np.arange(1, NE + 1)
    sgn = np.where(CD < 0, -1.0, 1.0)
    Fm, Fp = buf[:, :, K], buf[:, :, K + 1]
    Fn = np.where(sgn > 0, Fm, buf[:, :, K + 2])
    Fn1 = np.where(sgn > 0, buf[:, :, K - 1], Fp)
    FB = limiter_flux(Fm, Fp, Fn, Fn1, CD, CD / g.DE[None, None, :, None], sgn, beta)          # FBND(1..NE)
    WE = g.WE[None, None, 1:, None]
    new = Fl[:, :, 1:] - CD[:, :, 1:] / WE * FB[:, :, 1:] + CD[:, :, :-1] / WE * FB[:, :, :-1]
    new = np.where(new < 0, 1E-15, new)
    out = F.copy()
    out[1:, :, 1:] = new
    return out, Dt


def driftmu(g, inp, F2, S, DTs, beta):
    """DRIFTMU for species S (:382-473); returns (new F2 of the species, DtDriftMu)."""
    NR, NT, NE, NPA = g.NR, g.NT, g.NE, g.NPA
    s = S - 1
    QS = float(g.QS[s])
    F = F2[s].copy()
    BI, BH, B, VT, EIR, EIP, FH = inp.BOUNIS, inp.BOUNHS, inp.BNES, inp.VT, inp.EIR, inp.EIP, inp.FNHS
    RLZ = g.RLZ[:NR]
    I = np.arange(1, NR)
    J = np.arange(0, NT)
    J0 = np.where(J == 0, NT - 2, J - 1)
    J1 = np.where(J == NT - 1, 1, J + 1)
    L = np.arange(1, NPA)                      # Fortran L = 2..NPA
    ix = np.ix_
    R1 = RLZ[I][:, None]
    R3 = R1[:, :, None]
    DRM1 = (EIP[ix(I, J)] * R1 - (VT[ix(I, J1)] - VT[ix(I, J0)]) / 2 / g.DPHI) / B[ix(I, J)]
    DPM1 = OME * R1 + ((VT[ix(I + 1, J)] - VT[ix(I - 1, J)]) / 2 / g.MDR - EIR[ix(I, J)]) / B[ix(I, J)]
    MUBOUN = g.MU + 0.5 * g.WMU
    MUDOT = (1. - MUBOUN[None, :] ** 2) * DTs / 2 / MUBOUN[None, :] / RLZ[:, None]       # [I,L]
    MUDOT[:, NPA - 1] = 0.
    Bc = B[ix(I, J)][:, :, None]
    BIc, BHc = BI[ix(I, J, L)], BH[ix(I, J, L)]
    CMUDOT = MUDOT[ix(I, L)][:, None, :] * BIc / BHc                                   # [I,J,L]
    GMR1 = ((B[ix(I + 1, J)] - B[ix(I - 1, J)]) / 4 / g.MDR / B[ix(I, J)])[:, :, None]
    GMR2 = 1 / R3
    GMR3 = (BI[ix(I + 1, J, L)] - BI[ix(I - 1, J, L)]) / 2 / g.MDR / BIc
    GMP1 = ((B[ix(I, J1)] - B[ix(I, J0)]) / 4 / g.DPHI / B[ix(I, J)])[:, :, None]
    GMP2 = (BI[ix(I, J1, L)] - BI[ix(I, J0, L)]) / 2 / g.DPHI / BIc
    DRM2 = ((BI[ix(I, J1, L)] - BI[ix(I, J0, L)]) / 2 / g.DPHI
            + (BIc - 2 * BHc) * (B[ix(I, J1)] - B[ix(I, J0)])[:, :, None] / 4 / Bc / g.DPHI)
    DPM2 = (BIc + (BI[ix(I + 1, J, L)] - BI[ix(I - 1, J, L)]) * R3 / 2 / g.MDR
            + (BIc - 2 * BHc) * R3 / 4 / g.MDR * (B[ix(I + 1, J)] - B[ix(I - 1, J)])[:, :, None] / Bc)
    dBdt2 = (inp.dBdt[ix(I, J)] / 2. / B[ix(I, J)] * R1)[:, :, None]
    dIbndt2 = inp.dIbndt[ix(I, J, L)] * R3 / BIc
    eK = g.EKEV * 1e3 * (g.GREL[s] + 1) / 2 / g.GREL[s]                                  # [K]
    EDT = eK[None, None, :, None] / BHc[:, :, None, :] / R3[:, :, None, :] / Bc[:, :, None, :] / QS   # [I,J,K,L]
    DRDM = DRM1[:, :, None, None] + EDT * DRM2[:, :, None, :] * R3[:, :, None, :]
    DPDM = DPM1[:, :, None, None] - EDT * DPM2[:, :, None, :]
    CD = -CMUDOT[:, :, None, :] * ((GMR1 + GMR2 + GMR3)[:, :, None, :] * DRDM + (GMP1 + GMP2)[:, :, None, :] * DPDM
                                   + dBdt2[:, :, None, :] + dIbndt2[:, :, None, :])          # L = 2..NPA
    inside = (inp.outsideMGNP[ix(I, J)] == 0)
    ctemp = np.maximum(np.abs(CD), 1E-32)
    dts = np.where(inside[:, :, None, None], FRAC_CFL * DTs * g.DMU[None, None, None, 1:] / ctemp, np.inf)
    Dt = min(10000.0, float(dts.min()))
    Fl = F[1:].copy()                          # [I,J,K,L], line buffer with F(1) = F(2)
    Fl[:, :, :, 0] = Fl[:, :, :, 1]
    # limited flux for L = 2..NPA-2 (0-based 1..NPA-3); n = L+1-sgn
    Lf = np.arange(1, NPA - 2)
    c = CD[:, :, :, :NPA - 3]                  # CDriftMu at L = 2..NPA-2
    sgn = np.where(c < 0, -1.0, 1.0)
    Fm, Fp = Fl[..., Lf], Fl[..., Lf + 1]
    Fn = np.where(sgn > 0, Fm, Fl[..., Lf + 2])
    Fn1 = np.where(sgn > 0, Fl[..., Lf - 1], Fp)
    FBl = limiter_flux(Fm, Fp, Fn, Fn1, c, c / g.DMU[None, None, None, 1:NPA - 2], sgn, beta)
    # FBND(1) = 0, FBND(2..NPA-2) limited, FBND(NPA-1) = F(NPA); CDriftMu(.,1) = 0
    FB = np.concatenate([np.zeros_like(FBl[..., :1]), FBl, Fl[..., NPA - 1:NPA]], axis=3)     # index 0..NPA-2 = Fortran L = 1..NPA-1
    CDf = np.concatenate([np.zeros_like(CD[..., :1]), CD], axis=3)                          # index = Fortran L-1, L = 1..NPA
    Lu = np.arange(1, NPA - 1)                 # update L = 2..NPA-1
    WMU = g.WMU[None, None, None, Lu]
    new = F[1:][..., Lu] - CDf[..., Lu] / WMU * FB[..., Lu] + CDf[..., Lu - 1] / WMU * FB[..., Lu - 1]
    new = np.where(new < 0, 1E-15, new)
    out = F.copy()
    out[1:, :, :, 1:NPA - 1] = new
    out[1:, :, :, NPA - 1] = (out[1:, :, :, NPA - 2] * FH[ix(I, J)][:, :, None, NPA - 1] * g.MU[NPA - 1]
                              / FH[ix(I, J)][:, :, None, NPA - 2] / g.MU[NPA - 2])
    return out, Dt


def driftr(g, inp, F2, S, DTs, beta):
    """DRIFTR for species S (:95-198), line by line in the reference's (K, L, J) order because the
    ghost cells F(NR+1:NR+2) of the line buffer are only rewritten on inflow lines."""
    NR, NT, NE, NPA = g.NR, g.NT, g.NE, g.NPA
    s = S - 1
    QS = float(g.QS[s])
    F = F2[s].copy()
    FN, FH, B, VT, EIP = inp.FNIS, inp.FNHS, inp.BNES, inp.VT, inp.EIP
    RLZ = g.RLZ[:NR]
    I = np.arange(0, NR)
    J = np.arange(0, NT)
    J0 = np.where(J == 0, NT - 2, J - 1)
    J1 = np.where(J == NT - 1, 1, J + 1)
    ix = np.ix_
    VR = DTs / g.MDR / (RLZ + 0.5 * g.MDR) / 2 / g.DPHI
    sB = B[ix(I, J)] + B[ix(I + 1, J)]
    CR = (VR[:, None] * (VT[ix(I, J0)] + VT[ix(I + 1, J0)] - VT[ix(I, J1)] - VT[ix(I + 1, J1)]) / sB
          + (EIP[ix(I, J)] + EIP[ix(I + 1, J)]) / sB * DTs / g.MDR)                       # [I,J]
    CGR1 = FN[ix(I + 1, J1)] + FN[ix(I, J1)] - FN[ix(I + 1, J0)] - FN[ix(I, J0)]          # [I,J,L]
    CGR2 = (B[ix(I + 1, J1)] + B[ix(I, J1)] - B[ix(I + 1, J0)] - B[ix(I, J0)])[:, :, None]
    CGR3 = CGR1 + (FN[ix(I + 1, J)] + FN[ix(I, J)] - 2 * FH[ix(I + 1, J)] - 2 * FH[ix(I, J)]) * CGR2 / 2. / (B[ix(I + 1, J)] + B[ix(I, J)])[:, :, None]
    Dt = 100000.0
    inside = (inp.outsideMGNP == 0)
    buf = np.zeros(NR + 2)
    FB = np.zeros(NR)
    out = F.copy()
    for K in range(NE):
        P4 = DTs * g.EKEV[K] * 1000.0 * (g.GREL[s, K] + 1) / g.GREL[s, K] / g.DPHI / g.MDR / QS
        CGR = CGR3 / (FH[ix(I, J)] + FH[ix(I + 1, J)]) * P4 / 2. / sB[:, :, None] / (RLZ + 0.5 * g.MDR)[:, None, None]
        CD = CR[:, :, None] + CGR                                                         # [I,J,L]
        m = np.where(inside[:, :, None], FRAC_CFL * DTs / np.maximum(np.abs(CD), 1E-10), np.inf)
        Dt = min(Dt, float(m.min()))
        SG = np.where(CD < 0, -1, 1)
        for Lq in range(NPA):
            for j in range(NT):
                c, sg = CD[:, j, Lq], SG[:, j, Lq]
                buf[:NR] = F[:, j, K, Lq]
                if sg[NR - 1] == 1:
                    FB[0] = 0.
                    FB[NR - 1] = buf[NR - 1]
                    UR = NR - 1
                else:
                    FB[0] = buf[1]
                    UR = NR
                    if inp.outsideMGNP[NR - 1, j] == 1:
                        buf[NR] = 0.
                        buf[NR + 1] = 0.
                    else:
                        fg = inp.FGEOS[s, j, K, Lq]
                        buf[NR] = fg * g.CONF1 * FH[NR - 1, j, Lq]
                        buf[NR + 1] = fg * g.CONF2 * FH[NR - 1, j, Lq]
                ii = np.arange(1, UR)                                  # Fortran I = 2..UR
                sgi = sg[ii].astype(float)
                Fm, Fp = buf[ii], buf[ii + 1]
                Fn = np.where(sgi > 0, Fm, buf[np.minimum(ii + 2, NR + 1)])
                Fn1 = np.where(sgi > 0, buf[ii - 1], Fp)
                FB[ii] = limiter_flux(Fm, Fp, Fn, Fn1, c[ii], c[ii], sgi, beta)
                iu = np.arange(1, NR)
                new = buf[iu] - c[iu] * FB[iu] + c[iu - 1] * FB[iu - 1]
                out[1:, j, K, Lq] = np.where(new < 0, 1E-15, new)
    return out, Dt


# ---------------------------------------------------------------------------------------
# losses, moments, pitch-angle diffusion (src/ModRamLoss.f90:19-170, 457-507;
# src/ModRamRun.f90:231-259, 343-415; src/ModRamWPI.f90:643-714)
# exp / pow / log10 go through the C library one value at a time (numpy's vector loops may round
# the last bit differently from libm, which is what the C++ oracle calls).
# ---------------------------------------------------------------------------------------
import math

_exp = np.frompyfunc(math.exp, 1, 1)
_pow = np.frompyfunc(math.pow, 2, 1)

CEX_POLY = {0: (-18.767, -0.11017, -3.8173e-2, -0.1232, -5.0488e-2),        # Hydrogen
            2: (-20.789, 0.92316, -0.68017, 0.66153, -0.20998),            # HeliumP1
            1: (-18.987, -0.10613, -5.4841E-3, -1.6262E-2, -7.0554E-3)}    # OxygenP1


def cepara(g, inp, S, DTs):
    """CHARGE(I,J,K,L) of species S and ATLOS(I,K) (:19-170).  kind: 0 H+, 1 O+, 2 He+, 3 e- (no CEX)."""
    NR, NT, NE, NPA = g.NR, g.NT, g.NE, g.NPA
    s = S - 1
    CH = np.ones((NR, NT, NE, NPA))
    kind = int(g.kind[s])
    if kind in CEX_POLY:
        a0, a1, a2, a3, a4 = CEX_POLY[kind]
        for K in range(1, NE):
            X = math.log10(g.EKEV[K])
            if X < -2.:
                X = -2.
            # integer powers as the compiler expands them (powi): X**3 = X*X*X, X**4 = (X*X)*(X*X)
            Y = a0 + a1 * X + a2 * (X * X) + a3 * (X * X * X) + a4 * ((X * X) * (X * X))
            ALPHA = math.pow(10., Y) * g.V[s, K] * inp.HDNS[1:NR, :, 1:] * DTs
            CH[1:, :, K, 1:] = _exp(-ALPHA).astype(float)
    AT = np.zeros((NR, NE))
    for K in range(1, NE):
        for I in range(1, NR):
            TAUB = 2 * g.RLZ[I] / g.V[s, K]
            AT[I, K] = math.exp(-DTs / TAUB)
    return CH, AT


def charexchange(F, CH):
    out = F.copy()
    out[1:, :, 1:, 1:] = F[1:, :, 1:, 1:] * CH[1:, :, 1:, 1:]
    return out


def atmol(g, inp, F, AT):
    NR, NPA = g.NR, g.NPA
    out = F.copy()
    for I in range(1, NR):
        u = int(g.UPA[I])                       # Fortran L = u..NPA
        fac = _pow(AT[I, 1:][None, :, None], 1 / inp.FNHS[I, :, u - 1:][:, None, :]).astype(float)    # [J,K,L]
        out[I, :, 1:, u - 1:] = F[I, :, 1:, u - 1:] * fac
    return out


def sumrc(g, F):
    """SETRC in the reference's summation order I, K, L, J (:246-253)."""
    NR, NT, NE, NPA = g.NR, g.NT, g.NE, g.NPA
    W = (F[1:, :NT - 1, 1:, 1:] * g.WE[None, None, 1:, None] * g.WMU[None, None, None, 1:])
    T = g.EKEV[None, None, 1:, None] * W                 # [I,J,K,L]
    acc = 0.0
    for v in np.transpose(T, (0, 2, 3, 1)).ravel():      # I, K, L, J order
        acc = acc + v
    return acc


def anisch_pressures(g, inp, F, S):
    """PPERT, PPART(I,J) (:343-415) and the side effect F2(.,.,K,1) = F2(.,.,K,2); vectorised over
    (I,J), energy and pitch-angle sums in the reference's order."""
    NR, NT, NE, NPA = g.NR, g.NT, g.NE, g.NPA
    s = S - 1
    F = F.copy()
    RFAC = 4 * 3.1415926535897932384626433832795 / (CS * 100)
    PPERT = np.zeros((NR, NT))
    PPART = np.zeros((NR, NT))
    for I in range(1, NR):
        u = int(g.UPA[I] - 1)
        klo = 2
        for iwa in range(5):
            PPER = np.zeros(NT)
            PPAR = np.zeros(NT)
            for K in range(klo, int(g.khi[iwa]) + 1):       # Fortran K
                F[I, :, K - 1, 0] = F[I, :, K - 1, 1]
                SUME = np.zeros(NT)
                SUMA = np.zeros(NT)
                for L in range(1, u + 1):                   # Fortran L
                    ERNM = g.WMU[L - 1] / g.FFACTOR[s, I, K - 1, L - 1] / inp.FNHS[I, :, L - 1]
                    EPMA = ERNM * g.MU[L - 1] * g.MU[L - 1]
                    EPME = ERNM - EPMA
                    SUME = SUME + F[I, :, K - 1, L - 1] * EPME
                    SUMA = SUMA + F[I, :, K - 1, L - 1] * EPMA
                PPER = PPER + g.EPP[s, K - 1] * SUME
                PPAR = PPAR + g.EPP[s, K - 1] * SUMA
            PPAR = 2 * RFAC * PPAR
            PPER = RFAC * PPER
            klo = int(g.khi[iwa]) + 1
            PPERT[I] = PPERT[I] + PPER
            PPART[I] = PPART[I] + PPAR
    return PPERT, PPART, F


def wpadif(g, inp, F, DA, DB, DTs):
    """WPADIF with coefficient arrays DA + DB [I,J,K,L] (:643-714); vectorised over (I,J,K)."""
    NR, NT, NE, NPA = g.NR, g.NT, g.NE, g.NPA
    out = F.copy()
    Fl = F[1:, :, 1:, :]                                       # I = 2..NR, K = 2..NE
    FACMU = (inp.FNHS[1:NR] * g.MU[None, None, :])[:, :, None, :]   # [I,J,1,L]
    with np.errstate(divide="ignore", invalid="ignore"):
        f = Fl / FACMU                                         # L = 1 (MU = 0) is never used: F(1) = F(2)
    RK = np.zeros(Fl.shape)
    RL = np.zeros(Fl.shape)
    RL[..., 0] = -1.
    D = (DA + DB)[1:, :, 1:, :]
    for L in range(1, NPA - 1):                                # Fortran L = 2..NPA-1
        AN = D[..., L] / g.DMU[L]
        GN = D[..., L - 1] / g.DMU[L - 1]
        AN = AN * DTs / FACMU[..., L] / g.WMU[L]
        GN = GN * DTs / FACMU[..., L] / g.WMU[L]
        BN = AN + GN
        RP = f[..., L]
        DENOM = BN + GN * RL[..., L - 1] + 1
        RK[..., L] = (RP + GN * RK[..., L - 1]) / DENOM
        RL[..., L] = -AN / DENOM
    new = np.zeros(Fl.shape)
    new[..., NPA - 2] = RK[..., NPA - 2] / (1 + RL[..., NPA - 2])
    for L in range(NPA - 3, -1, -1):
        new[..., L] = RK[..., L] - RL[..., L] * new[..., L + 1]
    new[..., NPA - 1] = new[..., NPA - 2]
    out[1:, :, 1:, :] = new * FACMU
    return out


def wavelo(g, inp, F, DTs, use_plasmasphere=False):
    """WAVELO (src/ModRamWPI.f90:580-636): electron loss with the piecewise lifetime TAU_LIF."""
    NR, NT, NE, NPA = g.NR, g.NT, g.NE, g.NPA
    Bw = 100. if inp.Kp >= 4.0 else 30.
    RLpp = np.full(NT, 5.39 - 0.382 * inp.Kpmax12)
    if use_plasmasphere:
        for J in range(NT):
            for I in range(1, NR):
                if inp.NECR[I, J] > 50.:
                    RLpp[J] = g.LZ[I]
    out = F.copy()
    for K in range(1, NE):
        for I in range(1, NR):
            for J in range(NT):
                if g.LZ[I] <= RLpp[J]:
                    tau = inp.WALOS1[I, K] * ((10. / Bw) ** 2)
                else:
                    if g.EKEV[K] <= 1000.:
                        tau = inp.WALOS2[I, K] * (1 + inp.WALOS3[I, K] / inp.WALOS2[I, K])
                        if g.EKEV[K] <= 1.1:
                            tau = tau * 37.5813 * math.exp(-1.81255 * g.EKEV[K])
                        elif g.EKEV[K] <= 5.:
                            tau = tau * (7.5 - 1.15 * g.EKEV[K])
                    else:
                        tau = 5. * 3600 * 24 / inp.Kp
                out[I, J, K, 1:] = F[I, J, K, 1:] * math.exp(-DTs / tau)
    return out


# ---------------------------------------------------------------------------------------
# Coulomb collisions (src/ModRamCoul.f90:17-296)
# ---------------------------------------------------------------------------------------
PS_MASS = (5.4462E-4, 1.0, 4.0, 16.0, 14.0, 87.62)      # RAMSpecies(1:6): e-, H+, He+, O+, N+, Sr+
PS_CHARGE = (-1, 1, 1, 1, 1, 1)
PS_RATIO = (1.0, 0.77, 0.2, 0.03, 0.0, 0.0)
MP, RE_M, PI_R = 1.673E-27, 6.371E6, 3.1415926535897932384626433832795


def coulpara(g, S, DTs, gcoul):
    """COULE, COULI, ATA, GTA [K,L] of species S (:17-125).  The collision sums CCE.. are set to zero
    once, before the energy loop, and therefore accumulate over K (as written in the reference)."""
    NE, NPA = g.NE, g.NPA
    s = S - 1
    EPS, DLN = 8.854E-12, 21.5
    Zt = float(g.QS[s])
    QE = (Q ** 2 / EPS)
    GAMA = Zt ** 2 * DLN / 4. / PI_R * QE * 1E6 * QE
    CCO = GAMA / Q * DTs / Q / 1E3
    CCD = GAMA * DTs / (g.RMAS[s] * g.RMAS[s]) / (CS * CS * CS)
    COULE = np.zeros((NE, NPA)); COULI = np.zeros((NE, NPA)); ATA = np.zeros((NE, NPA)); GTA = np.zeros((NE, NPA))
    COULDE = np.zeros(NPA); COULDI = np.zeros(NPA)
    CCE = CDE = CCI = CDI = 0.0
    for K in range(NE):
        for b in range(6):
            RA = PS_RATIO[b]
            if RA < 1e-9:
                continue
            VF = math.sqrt(2. * Q / (MP * PS_MASS[b]))
            Zb = PS_CHARGE[b]
            X = g.VBND[s, K] / VF
            XD = g.V[s, K] / VF
            if Zb < 0:
                CCE = CCE + RA * gcoul(X)
                CDE = CDE + RA * (math.erf(XD) - gcoul(XD))
            else:
                CCI = CCI + RA * (Zb * Zb) * gcoul(X)
                CDI = CDI + RA * (Zb * Zb) * (math.erf(XD) - gcoul(XD))
        COULE[K, 0] = -CCE * g.VBND[s, K] * CCO * (g.GRBND[s, K] * g.GRBND[s, K])
        COULI[K, 0] = -CCI * g.VBND[s, K] * CCO * (g.GRBND[s, K] * g.GRBND[s, K])
        CCDE = CCD * CDE * g.GREL[s, K] / math.pow(g.GREL[s, K] * g.GREL[s, K] - 1, 1.5)
        CCDI = CCD * CDI * g.GREL[s, K] / math.pow(g.GREL[s, K] * g.GREL[s, K] - 1, 1.5)
        for L in range(1, NPA - 1):            # Fortran L = 2..NPA-1
            COULE[K, L] = COULE[K, 0]
            COULI[K, L] = COULI[K, 0]
            MUBOUN = g.MU[L] + 0.5 * g.WMU[L]
            BADIF = (1. - MUBOUN * MUBOUN) / MUBOUN / 2.
            COULDE[L] = CCDE * BADIF
            AFER = COULDE[L] / g.MU[L] / g.DMU[L] / g.WMU[L]
            ASEC = COULDE[L - 1] / g.MU[L] / g.DMU[L - 1] / g.WMU[L]
            COULDI[L] = CCDI * BADIF
            AFIR = COULDI[L] / g.MU[L] / g.DMU[L] / g.WMU[L]
            ASIC = COULDI[L - 1] / g.MU[L] / g.DMU[L - 1] / g.WMU[L]
            ATA[K, L] = AFIR + AFER
            GTA[K, L] = ASIC + ASEC
        ATA[K, NPA - 1] = 0
    return COULE, COULI, ATA, GTA


def coulen(g, inp, F, S, COULE, COULI, beta):
    """COULEN (:133-221): energy drag, the limiter of DRIFTE on CccolE = (COULE+COULI)*NECR*BANE(L);
    vectorised over (I, J, L)."""
    NR, NT, NE, NPA = g.NR, g.NT, g.NE, g.NPA
    s = S - 1
    out = F.copy()
    Fl = F[1:]                                                   # [I,J,K,L]
    Ls = np.arange(1, NPA)                                       # Fortran L = 2..NPA
    Lb = np.minimum(Ls, NPA - 12)                                # BANE(L >= NPA-10) = BANE(NPA-11)  (0-based NPA-12)
    BANE = (1. - inp.FNIS[1:NR][:, :, Lb] / 2. / inp.FNHS[1:NR][:, :, Lb]) / (1. - g.MU[Lb] * g.MU[Lb])[None, None, :]
    XNE = inp.NECR[1:NR][:, :, None] * BANE                      # [I,J,L]
    G = g.GREL[s]
    EZERO = g.EKEV[0] - g.WE[0]
    GRZ = 1. + EZERO * 1000. * Q / g.RMAS[s] / CS / CS
    Fs = Fl[:, :, :, 1:]                                         # L = 2..NPA
    f1 = Fs[:, :, 1] * G[0] / G[1] * np.sqrt((G[0] ** 2 - 1) / (G[1] ** 2 - 1))
    f0 = f1 * GRZ / G[0] * np.sqrt((GRZ ** 2 - 1) / (G[0] ** 2 - 1))
    z = np.zeros_like(f1)
    buf = np.concatenate([f0[:, :, None], f1[:, :, None], Fs[:, :, 1:], z[:, :, None], z[:, :, None]], axis=2)   # index = Fortran K
    CD = (COULE[:, 1:] + COULI[:, 1:])[None, None, :, :] * XNE[:, :, None, :]           # [I,J,K,L]
    K = np.arange(1, NE + 1)
    sgn = np.where(CD < 0, -1.0, 1.0)
    Fm, Fp = buf[:, :, K], buf[:, :, K + 1]
    Fn = np.where(sgn > 0, Fm, buf[:, :, K + 2])
    Fn1 = np.where(sgn > 0, buf[:, :, K - 1], Fp)
    FB = limiter_flux(Fm, Fp, Fn, Fn1, CD, CD / g.DE[None, None, :, None], sgn, beta)
    WE = g.WE[None, None, 1:, None]
    new = Fs[:, :, 1:] - CD[:, :, 1:] / WE * FB[:, :, 1:] + CD[:, :, :-1] / WE * FB[:, :, :-1]
    new = np.where(new < 0, 1E-15, new)
    out[1:, :, 1:, 1:] = new
    return out


def coulmu(g, inp, F, S, ATA, GTA, T):
    """COULMU (:229-296): implicit pitch-angle scattering; vectorised over (I, J, K)."""
    NR, NT, NE, NPA = g.NR, g.NT, g.NE, g.NPA
    out = F.copy()
    Fl = F[1:, :, 1:, :]                                         # I = 2..NR, K = 2..NE
    XNE = inp.NECR[1:NR][:, :, None]
    BI, BH, FH = inp.BOUNIS[1:NR], inp.BOUNHS[1:NR], inp.FNHS[1:NR]
    BAS = XNE * BI / 2. / BH                                     # BASCNE(I,J,L)
    RK = np.zeros(Fl.shape)
    RL = np.zeros(Fl.shape)
    RL[..., 0] = -1.
    for L in range(1, NPA - 1):
        AN = ATA[None, None, 1:, L] * BAS[:, :, None, L] / FH[:, :, None, L] * BH[:, :, None, L]
        GN = GTA[None, None, 1:, L] * BAS[:, :, None, L - 1] / FH[:, :, None, L] * BH[:, :, None, L - 1]
        BN = AN + GN
        RP = Fl[..., L] / FH[:, :, None, L] / g.MU[L]
        DENOM = BN + GN * RL[..., L - 1] + 1
        RK[..., L] = (RP + GN * RK[..., L - 1]) / DENOM
        RL[..., L] = -AN / DENOM
    new = np.zeros(Fl.shape)
    new[..., NPA - 2] = RK[..., NPA - 2] / (1 + RL[..., NPA - 2])
    for L in range(NPA - 3, -1, -1):
        new[..., L] = RK[..., L] - RL[..., L] * new[..., L + 1]
    new[..., NPA - 1] = new[..., NPA - 2]
    new = new * FH[:, :, None, :] * g.MU[None, None, None, :]
    if T > 0:
        new = np.where(new < 0, 1E-15, new)
    out[1:, :, 1:, :] = new
    return out


# ---------------------------------------------------------------------------------------
# the species loop and epilogue of ram_run (src/ModRamRun.f90:64-222)
# ---------------------------------------------------------------------------------------
def para_flc(g, inp, S, r_curvEq, zeta1Eq, zeta2Eq):
    """PARA_FLC (src/ModRamLoss.f90:342-455), whole-array numpy: FLC_coef (NR,NT,NE,NPA) of species S (1-based)."""
    Q, REarth = 1.602E-19, 6.4 * 1.E6
    NR, NT, NE, NPA = g.NR, g.NT, g.NE, g.NPA
    V = g.V[S - 1][None, None, :]                                     # (1,1,NE)
    BNES = inp.BNES[:NR, :, None]
    rg = g.RMAS[S - 1] * V / np.abs(BNES * Q)
    eps = np.minimum(rg / r_curvEq[:, :, None], 0.584)
    on = eps >= 0.1
    e = np.where(on, eps, 0.3)                                        # placeholder where the coefficient stays 0
    e1, e2, e3 = 1.0 / e, 1.0 / (e * e), 1.0 / (e * e * e)
    a1 = -0.35533865 + 0.12800347 * e1 + 0.0017113113 * e2
    a2 = 0.23156321 + 0.15561211 * e1 - 0.001860433 * e2
    ba = -0.51057275 + 0.93651781 * e1 - 0.031690658 * e2
    ca = 1.0663037 - 1.0944973 * e1 + 0.016679378 * e2 - 0.000499 * e3
    da = -0.49667826 - 0.0081941799 * e1 + 0.0013621659 * e2
    om = 1.0513540 + 0.1351358 * e - 0.50787555 * (e * e)
    Am = np.exp(ca) * (zeta1Eq[:, :, None] ** a1 * zeta2Eq[:, :, None] ** a2 + da)
    mub = (g.MU + 0.5 * g.WMU)[:NPA - 1]                              # (NPA-1,)
    alph = np.arccos(mub)
    sn = np.sin(om[..., None] * alph)                                 # (NR,NT,NE,NPA-1)
    pw = mub ** ba[..., None]
    Nf = 1.0 / (sn * pw)
    lmin = (Nf.shape[-1] - 1) - np.argmin(Nf[..., ::-1], axis=-1)     # `<=` keeps the LAST minimum
    nfm = np.take_along_axis(Nf, lmin[..., None], axis=-1)
    bh = inp.BOUNHS[:NR, :, None, :NPA - 1]
    tau = 4 * g.LZ[:NR, None, None, None] * REarth * bh / V[..., None]
    D = (Am * Am)[..., None] / (2 * tau)
    Daa = D * (nfm * nfm) * (sn * sn) * mub ** (2 * ba[..., None]) / ((1 - mub * mub) * (mub * mub))
    out = np.zeros((NR, NT, NE, NPA), order="F")
    out[..., :NPA - 1] = np.where(on[..., None], Daa * (1 - mub * mub) * mub * bh, 0.0)
    return out


def ram_run(g, inp, F2, DTs, beta, gcoul, DtsMin=1.0, T=0.0, wpi=False, emic=False, coulomb=False, DAA=None):
    """One ram_run step of all species with the restatements above, in the reference's call order.
    Returns F2, DtsNext, SETRC per species, the loss increments LSDR/LSCHA/LSATM/LSWAE/LSCOE/LSCSC and
    PPERT/PPART.  DAA: (ATAW+ATAC for electrons, ATAW_emic_h+ATAW_emic_he for H+) or None."""
    nS, NR, NT = g.nS, g.NR, g.NT
    F2 = F2.copy()
    dts = {n: np.zeros(nS) for n in ("R", "P", "E", "MU")}
    loss = {n: np.zeros(nS) for n in ("DR", "CHA", "ATM", "WAE", "COE", "CSC")}
    SETRC = np.zeros(nS)
    for S in range(1, nS + 1):
        s = S - 1
        kind = int(g.kind[s])
        sWPI, sCEX, sEMIC = kind == 3, kind != 3, kind == 0
        CH, AT = cepara(g, inp, S, DTs)
        CE, CI, ATA, GTA = coulpara(g, S, DTs, gcoul)
        F = F2[s]
        setrc = 0.0                     # SETRC(S) before the first SUMRC of the run

        def SUM(key):
            nonlocal setrc
            new = sumrc(g, F2[s])
            loss[key][s] += setrc - new
            setrc = new

        def put(Fnew):
            F2[s] = Fnew

        for fn, key in ((driftr, "R"), (driftp, "P"), (drifte, "E"), (driftmu, "MU")):
            Fn, dt = fn(g, inp, F2, S, DTs, beta)
            put(Fn)
            dts[key][s] = dt
        SUM("DR")
        if coulomb:
            put(coulen(g, inp, F2[s], S, CE, CI, beta)); SUM("COE")
            put(coulmu(g, inp, F2[s], S, ATA, GTA, T)); SUM("CSC")

        def waves():
            if sWPI:
                put(wpadif(g, inp, F2[s], DAA[0], DAA[1], DTs) if wpi else wavelo(g, inp, F2[s], DTs)); SUM("WAE")

        def emicw():
            if sEMIC and emic:
                put(wpadif(g, inp, F2[s], DAA[2], DAA[3], DTs)); SUM("WAE")

        def cex():
            if sCEX:
                put(charexchange(F2[s], CH)); SUM("CHA")

        waves(); emicw(); cex()
        put(atmol(g, inp, F2[s], AT)); SUM("ATM")
        put(atmol(g, inp, F2[s], AT)); SUM("ATM")
        cex(); emicw(); waves()
        if coulomb:
            put(coulmu(g, inp, F2[s], S, ATA, GTA, T)); SUM("CSC")
            put(coulen(g, inp, F2[s], S, CE, CI, beta)); SUM("COE")
        for fn, key in ((driftmu, "MU"), (drifte, "E"), (driftp, "P"), (driftr, "R")):
            Fn, dt = fn(g, inp, F2, S, DTs, beta)
            put(Fn)
            dts[key][s] = dt
        SUM("DR")
        SETRC[s] = setrc
    F2[:, :, NT - 1] = F2[:, :, 0]
    out = inp.outsideMGNP == 1
    F2[:, out] = 1.e-31
    DtsNext = max(min(dts["R"].min(), dts["P"].min(), dts["E"].min(), dts["MU"].min()), DtsMin)
    PPERT = np.zeros((nS, NR, NT)); PPART = np.zeros((nS, NR, NT))
    for S in range(1, nS + 1):
        pe, pa, Fn = anisch_pressures(g, inp, F2[S - 1], S)
        F2[S - 1] = Fn
        PPERT[S - 1], PPART[S - 1] = pe, pa
    return F2, DtsNext, SETRC, loss, PPERT, PPART


# ---------------------------------------------------------------------------------------------
# ANISCH, second half: the diffusion-coefficient rebuild (src/ModRamRun.f90:422-605), restated a second time:
# whole-array numpy, the Steffen spline through tests/independent_scb.interp1d_steffen (searchsorted instead of the
# bisection loop), the bilinear rule written on index arrays.
def anisch_diffcoef(g, inp, t, S, wpi, emic, Kp, AE=0, use_bas=True):
    from independent_scb import interp1d_steffen
    NR, NT, NE, NPA = g.NR, g.NT, g.NE, g.NPA
    s = S - 1
    CSv, PI = 2.998E8, 3.1415926535897932384626433832795
    MUB = g.MU + 0.5 * g.WMU
    XNE, B, BH = t["XNE"], inp.BNES, inp.BOUNHS
    out = {}

    def bilinear(xa, ya, za, x, y):                     # za[i, j] = f(xa[i], ya[j]); x, y arrays of equal shape
        xi = np.clip(np.searchsorted(xa, x, side="right") - 1, 0, len(xa) - 2)
        yi = np.clip(np.searchsorted(ya, y, side="right") - 1, 0, len(ya) - 2)
        tt = (x - xa[xi]) / (xa[xi + 1] - xa[xi])
        u = (y - ya[yi]) / (ya[yi + 1] - ya[yi])
        return ((1. - tt) * (1. - u) * za[xi, yi] + tt * (1. - u) * za[xi + 1, yi] + (1. - tt) * u * za[xi, yi + 1]
                + tt * u * za[xi + 1, yi + 1])

    if wpi and g.kind[s] == 3:
        ATAW = np.zeros((NR, NT, NE, NPA), order="F")
        ATAC = np.zeros((NR, NT, NE, NPA), order="F")
        PA = 180.0 / PI * np.arccos(g.MU[::-1])
        tab = t["CDAAR"][..., ::-1] if use_bas else t["BDAAR"]
        Bw = 100. if Kp >= 4.0 else 30.
        esu, cv, gausgam = 1.602E-19 * 3E9, CSv * 100, 1.E-5
        ALENOR = np.log10(t["ENOR"])
        ER1 = np.log10(g.EKEV[1:])
        for I in range(1, NR):
            for J in range(NT):
                if XNE[I, J] <= 50.:
                    for K in range(1, NE):
                        Y = interp1d_steffen(PA, np.log10(tab[I, J, K, :]), t["PAbn"])
                        tau = 10. ** Y * 1
                        tau = np.where(tau > 1e0, 1e-1, tau)
                        tau = np.where(tau < 1e-30, 1e-30, tau)
                        ATAC[I, J, K, :] = tau * (1. - MUB * MUB) * MUB * BH[I, J, :]
                else:
                    omega = esu * 10 * B[I, J] / (g.RMAS[s] * cv)
                    xfrl = min(max(CSv * np.sqrt(XNE[I, J] * g.RMAS[s] * 40 * PI) / 10. / B[I, J], 2.), 18.)
                    fnorm = omega * ((Bw * 1e-3) * (Bw * 1e-3)) * (gausgam * gausgam) / 1e8 / B[I, J] / B[I, J]
                    for L in range(NPA):
                        Y = bilinear(ALENOR, t["fpofc"], np.log10(t["NDAAJ"][I, :, L, :]), ER1, np.full_like(ER1, xfrl))
                        ATAW[I, J, 1:, L] = 10. ** Y * fnorm / (g.GREL[s, 1:] * g.GREL[s, 1:]) * (1. - MUB[L] * MUB[L]) / MUB[L]
        out["ATAW"], out["ATAC"] = ATAW, ATAC
    if emic and g.kind[s] == 0:
        AH = np.zeros((NR, NT, NE, NPA), order="F")
        AHE = np.zeros((NR, NT, NE, NPA), order="F")
        cls = 1 if 0 <= AE < 100 else (2 if 100 <= AE < 300 else (3 if 300 <= AE < 400 else (4 if AE >= 400 else 0)))
        me = g.RMAS[int(np.argmax(g.kind == 3))]
        logE = np.log10(t["EKEV_emic"])
        ER1 = np.log10(g.EKEV[1:])
        for I in range(1, NR):
            for J in range(NT):
                xfrl = min(max(CSv * np.sqrt(XNE[I, J] * me * 40 * PI) / 10. / B[I, J], 2.), 20.)
                fh = t["Ihs_emic"][cls - 1, I, J] if cls else 0.0
                fhe = t["Ihes_emic"][cls - 1, I, J] if cls else 0.0
                for L in range(NPA):
                    for arr, tab, f in ((AH, t["Daa_emic_h"], fh), (AHE, t["Daa_emic_he"], fhe)):
                        Y = bilinear(logE, t["fp2c_emic"], np.log10(tab[I, :, L, :]), ER1, np.full_like(ER1, xfrl))
                        v = 10. ** Y * f * (1. - MUB[L] * MUB[L]) * MUB[L] * BH[I, J, L]
                        arr[I, J, 1:, L] = np.where(v <= 1.0e-20, 1.0e-31, v)
        out["ATAW_emic_h"], out["ATAW_emic_he"] = AH, AHE
    return out
