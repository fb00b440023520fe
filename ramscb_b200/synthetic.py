"""Seeded synthetic inputs for the RAM hot path.

The reference's real inputs (initialization.nc, boundary-flux and index files)
are missing blobs, and its field arrays come from the SCB coupling
(``computehI``, src/ModRamScb.f90:568-626).  These generators produce inputs of
the same *shape and magnitude*: an Ejiri dipole (src/ModRamFunctions.f90:90-143)
with a smooth day-night perturbation so every azimuthal-gradient term of the
drift coefficients is exercised, a Volland-Stern potential
(src/ModRamRun.f90:45-51), a Rairden geocorona (src/ModRamScb.f90:365-370), and
Maxwellian / log-normal-noise / adversarial phase-space distributions.

All arrays are Fortran-ordered float64 with the reference's shapes
(src/ModRamInit.f90:68-151).  Both the CUDA library and the CPU oracle consume
the same bytes.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from . import grids as G


def _f(shape):
    return np.zeros(shape, dtype=np.float64, order="F")


@dataclass
class RamInputs:
    BNES: np.ndarray      # (NR+1,NT)
    dBdt: np.ndarray
    VT: np.ndarray
    EIR: np.ndarray
    EIP: np.ndarray
    FNHS: np.ndarray      # (NR+1,NT,NPA)
    FNIS: np.ndarray
    BOUNHS: np.ndarray
    BOUNIS: np.ndarray
    HDNS: np.ndarray
    dIdt: np.ndarray
    dIbndt: np.ndarray
    outsideMGNP: np.ndarray  # int32 (NR,NT)
    NECR: np.ndarray      # (NR,NT)
    FGEOS: np.ndarray     # (nS,NT,NE,NPA)
    F2: np.ndarray        # (nS,NR,NT,NE,NPA)
    Kp: float
    Kpmax12: float
    WALOS1: np.ndarray = None  # (NR,NE) for the WPI species
    WALOS2: np.ndarray = None
    WALOS3: np.ndarray = None


def volland_stern(g: G.RamGrids, Kp: float, PHIOFS: float = 0.0):
    """VT(NR+1,NT) -- src/ModRamRun.f90:45-51."""
    AVS = 7.05e-6 / (1.0 - 0.159 * Kp + 0.0093 * Kp ** 2) ** 3 / G.RE
    VT = _f((g.NR + 1, g.NT))
    for I in range(g.NR + 1):
        for J in range(g.NT):
            VT[I, J] = AVS * (g.LZ[I] * G.RE) ** 2 * math.sin(g.PHI[J] - PHIOFS)
    return VT


def make_inputs(g: G.RamGrids, seed: int = 20240317, f2_kind: str = "noisy",
                eps: float = 0.1, inductive: bool = False, efield_ind: bool = False,
                mgnp: bool = False, Kp: float = 3.0) -> RamInputs:
    """Build one complete, self-consistent input set.

    f2_kind: "smooth" (Maxwellian x sin^n), "noisy" (x log-normal noise, hits
    every limiter branch), "adversarial" (plateaus, 1e-15 cells, sign-alternating
    slopes).
    """
    nS, NR, NT, NE, NPA = g.nS, g.NR, g.NT, g.NE, g.NPA
    rng = np.random.default_rng(seed)
    LZ = g.LZ                                    # (NR+1)
    cosphi = np.cos(g.PHI)
    sinphi = np.sin(g.PHI)
    cosphi[-1], sinphi[-1] = cosphi[0], sinphi[0]  # J=NT is the J=1 meridian

    # -- equatorial field, compressed on the day side ----------------------------
    BNES = _f((NR + 1, NT))
    BNES[:, :] = (0.32 / LZ ** 3 / 1.0e4)[:, None] * (1.0 + eps * (LZ[:, None] / 6.5) ** 2 * cosphi[None, :])

    # -- bounce integrals h, I at cell centres and cell boundaries ---------------
    hC = np.array([G.funt(m) for m in g.MU])
    iC = np.array([G.funi(m) for m in g.MU])
    mub = np.minimum(g.MU + 0.5 * g.WMU, 1.0)
    hB = np.array([G.funt(m) for m in mub])
    iB = np.array([G.funi(m) for m in mub])
    pert_h = 1.0 + 0.5 * eps * (LZ[:, None, None] / 6.5) ** 2 * cosphi[None, :, None] * (1.0 - 0.3 * g.MU[None, None, :])
    pert_i = 1.0 + 0.7 * eps * (LZ[:, None, None] / 6.5) ** 2 * cosphi[None, :, None] * (1.0 + 0.2 * g.MU[None, None, :])
    FNHS = np.asfortranarray(hC[None, None, :] * pert_h)
    FNIS = np.asfortranarray(iC[None, None, :] * pert_i)
    BOUNHS = np.asfortranarray(hB[None, None, :] * pert_h)
    BOUNIS = np.asfortranarray(iB[None, None, :] * pert_i)

    # -- geocorona: Rairden profile, a little denser for field-aligned particles --
    r = LZ
    dens = 10.0 ** (13.326 - 3.6908 * r + 1.1362 * r ** 2 - 0.16984 * r ** 3 + 0.009553 * r ** 4)
    HDNS = np.asfortranarray(dens[:, None, None] * (1.0 + 0.5 * g.MU[None, None, :] ** 2)
                             * (1.0 + 0.0 * cosphi[None, :, None]))

    dBdt = _f((NR + 1, NT))
    dIdt = _f((NR + 1, NT, NPA))
    dIbndt = _f((NR + 1, NT, NPA))
    if inductive:
        # a slow compression: dB/dt ~ 1e-3 B per second on the day side, I shrinking
        dBdt[:, :] = 1.0e-3 * BNES * cosphi[None, :]
        dIdt[:, :, :] = -5.0e-4 * FNIS * (0.5 + 0.5 * cosphi[None, :, None])
        dIbndt[:, :, :] = -5.0e-4 * BOUNIS * (0.5 + 0.5 * cosphi[None, :, None])
        dBdt[0, :] = 0.0
        dIdt[0, :, :] = 0.0
        dIbndt[0, :, :] = 0.0

    VT = volland_stern(g, Kp)
    VT[:, -1] = VT[:, 0]
    EIR = _f((NR + 1, NT))
    EIP = _f((NR + 1, NT))
    if efield_ind:
        EIR[1:, :] = 2.0e-5 * sinphi[None, :] * (LZ[1:, None] / 6.5)
        EIP[1:, :] = -1.5e-5 * cosphi[None, :] * (LZ[1:, None] / 6.5)

    outside = np.zeros((NR, NT), dtype=np.int32, order="F")
    if mgnp:
        # a few dayside cells of the outermost shells are flagged as outside the
        # magnetopause (src/ModRamRun.f90:197-201, DRIFTR :161-163)
        jn = NT // 2
        outside[NR - 1, jn - 1:jn + 2] = 1
        outside[NR - 2, jn] = 1

    NECR = _f((NR, NT))
    NECR[:, :] = (10.0 ** (-0.3145 * LZ[:NR] + 3.9043))[:, None]

    # -- phase-space density -------------------------------------------------------
    # differential flux j(E,alpha,L) [1/cm2/s/sr/keV], F2 = j * FFACTOR * FNHS
    kT = np.array([5.0 if sp.charge > 0 else 1.0 for sp in g.species])
    amp = np.array([3e5, 8e4, 2e4, 6e6])[:nS]
    E = g.EKEV
    sina = np.sqrt(np.maximum(1.0 - g.MU ** 2, 0.0))
    F2 = _f((nS, NR, NT, NE, NPA))
    FGEOS = _f((nS, NT, NE, NPA))
    u_out = int(g.UPA[NR - 1])
    for s in range(nS):
        spec = amp[s] * (E / kT[s]) * np.exp(-E / kT[s]) + amp[s] * 1e-3 * (1.0 + E / 20.0) ** (-3.5)
        pa = 0.15 + 0.85 * sina                                  # sin^1 anisotropy with a floor
        rad = np.exp(-((LZ[:NR] - 4.0) / 1.2) ** 2) + 0.05
        mlt = 1.0 + 0.3 * cosphi - 0.2 * sinphi
        j = (rad[:, None, None, None] * mlt[None, :, None, None]
             * spec[None, None, :, None] * pa[None, None, None, :])
        f = j * g.FFACTOR[s][:, None, :, :] * FNHS[:NR, :, None, :]
        # empty loss cone
        for I in range(NR):
            f[I, :, :, int(g.UPA[I]) - 1:] *= 1e-4
        F2[s] = f
        # boundary flux: kappa-like spectrum, isotropic outside the loss cone
        kap = amp[s] * 0.3 * (1.0 + E / (3.0 * kT[s])) ** (-4.0)
        fg = (mlt[:, None, None] * kap[None, :, None] * np.ones(NPA)[None, None, :]
              * g.FFACTOR[s][NR - 1][None, :, :])
        fg[:, :, u_out - 1:] = 0.0
        FGEOS[s] = fg

    if f2_kind in ("noisy", "adversarial"):
        F2 *= np.asfortranarray(np.exp(rng.standard_normal(F2.shape)))
        FGEOS *= np.asfortranarray(np.exp(0.5 * rng.standard_normal(FGEOS.shape)))
    if f2_kind == "adversarial":
        # plateaus (|dF| <= 1e-27), cells at the clamp value, tiny cells
        m = rng.random(F2.shape) < 0.05
        F2[m] = 1.0e-15
        for s in range(nS):
            F2[s, 3:6, :, 5:9, 10:14] = 7.25                     # exact plateau in all 4 dims
            F2[s, :, :, 12, :] = F2[s, :, :, 11, :]               # plateau pair along E
            F2[s, 8, :, :, :] = F2[s, 7, :, :, :]                 # plateau pair along R
        F2[:, :, :, :, 30] *= 1.0e-12                             # deep trough along mu
        F2[:, :, 7, :, :] *= 1.0e+6                               # wall along phi
    F2[:, :, NT - 1, :, :] = F2[:, :, 0, :, :]
    FGEOS[:, NT - 1, :, :] = FGEOS[:, 0, :, :]

    inp = RamInputs(BNES=BNES, dBdt=dBdt, VT=VT, EIR=EIR, EIP=EIP, FNHS=FNHS, FNIS=FNIS,
                    BOUNHS=BOUNHS, BOUNIS=BOUNIS, HDNS=HDNS, dIdt=dIdt, dIbndt=dIbndt,
                    outsideMGNP=outside, NECR=NECR, FGEOS=FGEOS, F2=F2, Kp=Kp, Kpmax12=Kp)
    for s, sp in enumerate(g.species):
        if sp.WPI:
            inp.WALOS1, inp.WALOS2, inp.WALOS3 = G.wavepara(g, s)
    if inp.WALOS1 is None:
        inp.WALOS1 = _f((NR, NE)) + 1.0
        inp.WALOS2 = _f((NR, NE)) + 1.0
        inp.WALOS3 = _f((NR, NE)) + 1.0
    return inp


def synthetic_daa(g: G.RamGrids, inp: RamInputs):
    """Synthetic pitch-angle diffusion coefficient D(NR,NT,NE,NPA) for WPADIF.

    The tabulated Daa files are missing blobs; this keeps the functional wrapper
    and clipping the reference applies to its look-ups
    (src/ModRamRun.f90:458-467 chorus, :574-585 EMIC):
    D = tau * (1-mub^2) * mub * BOUNHS, tau = 1e-4 (E/10 keV)^-1/2 1/s clipped
    to [1e-30, 1e-1].
    """
    NR, NT, NE, NPA = g.NR, g.NT, g.NE, g.NPA
    mub = np.minimum(g.MU + 0.5 * g.WMU, 1.0)
    tau = np.clip(1.0e-4 * (g.EKEV / 10.0) ** -0.5, 1e-30, 1e-1)
    D = (tau[None, None, :, None] * ((1.0 - mub ** 2) * mub)[None, None, None, :]
         * inp.BOUNHS[:NR, :, None, :])
    return np.asfortranarray(np.clip(D, 1e-30, None))


def synthetic_wave_tables(g: G.RamGrids, inp: RamInputs, seed: int = 7):
    """Synthetic stand-ins for the tabulated bounce-averaged diffusion coefficients the reference reads at start-up
    (src/ModRamWPI.f90:185-470; the files are missing blobs) -- same shapes, axes and orders of magnitude -- and the
    plasmaspheric density XNE with a plasmapause near L = 4.5, so that both branches of the ANISCH rebuild
    (src/ModRamRun.f90:447-515: chorus outside, hiss inside) are taken.  Returns a dict of Fortran-ordered arrays."""
    rng = np.random.default_rng(seed)
    NR, NT, NE, NPA = g.NR, g.NT, g.NE, g.NPA
    ENG, NCF, ENGe, NCFe = 45, 5, 41, 10                       # src/ModRamGrids.f90:31-34, :51
    t = {"ENG": ENG, "NCF": NCF, "ENG_emic": ENGe, "NCF_emic": NCFe}
    t["ENOR"] = np.logspace(-2.0, 3.5, ENG)                     # normalised energies, ascending
    t["fpofc"] = 2.0 + 4.0 * np.arange(NCF)                     # src/ModRamWPI.f90:351-362
    t["EKEV_emic"] = np.logspace(-1.0, 3.0, ENGe)               # 0.1 keV .. 1000 keV
    t["fp2c_emic"] = 2.0 * (1 + np.arange(NCFe))                # :232
    mu = g.MU
    shape_l = 0.2 + np.sin(np.pi * np.clip(mu, 0.02, 0.98)) ** 2

    def smooth(shape, lo, hi):
        return np.asfortranarray(10.0 ** (lo + (hi - lo) * rng.random(shape)))

    t["NDAAJ"] = np.asfortranarray(smooth((NR, ENG, NPA, NCF), -3.0, 1.0) * shape_l[None, None, :, None])
    t["CDAAR"] = np.asfortranarray(smooth((NR, NT, NE, NPA), -7.0, -3.0) * shape_l[None, None, None, :])
    t["BDAAR"] = np.asfortranarray(smooth((NR, NT, NE, NPA), -7.0, -3.0) * shape_l[None, None, None, :])
    t["Daa_emic_h"] = np.asfortranarray(smooth((NR, ENGe, NPA, NCFe), -8.0, -2.0) * shape_l[None, None, :, None])
    t["Daa_emic_he"] = np.asfortranarray(smooth((NR, ENGe, NPA, NCFe), -9.0, -3.0) * shape_l[None, None, :, None])
    t["Daa_emic_h"][:, :, ::7, :] *= 1e-30                      # some entries below the 1e-20 floor (:585-586)
    t["Ihs_emic"] = np.asfortranarray(0.1 + rng.random((4, NR, NT)))
    t["Ihes_emic"] = np.asfortranarray(0.05 + rng.random((4, NR, NT)))
    LZ = g.LZ[:NR]
    XNE = np.asfortranarray(inp.NECR[:NR, :NT] * np.where(LZ[:, None] > 4.5, 0.05, 1.0))
    # J = NT is the J = 1 meridian (PHI = 2 pi): every (I,J) input of the reference repeats there, and the step relies on it
    # (F2(J=1) = F2(J=NT) is maintained by DRIFTP, src/ModRamDrift.f90:272, and by ram_run's epilogue)
    XNE[:, -1] = XNE[:, 0]
    for n in ("CDAAR", "BDAAR"):
        t[n][:, -1] = t[n][:, 0]
    for n in ("Ihs_emic", "Ihes_emic"):
        t[n][:, :, -1] = t[n][:, :, 0]
    t["XNE"] = XNE
    t["PAbn"] = np.asfortranarray(g.PAbn, dtype=np.float64)
    return t
