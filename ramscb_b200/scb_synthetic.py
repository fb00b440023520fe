"""Seeded synthetic inputs for the SCB hot path.

Host-side setup that the reference does in ``computational_domain`` (DIPS branch,
src/ModScbIO.f90:71-82,179-205), ``get_dipole_lines`` (src/ModScbFunctions.f90:13-54),
``scb_init`` (src/ModScbInit.f90:230-290) and the tail of ``pressure``
(src/ModScbRun.f90:1107-1175): dipole field-line geometry, Euler-potential
ladders, an analytic anisotropic pressure mapped along the field lines with the
iLossCone=1 formulas, and its Steffen-spline derivatives.  None of this is on
the GPU hot path; it only produces arrays of the right shape and magnitude that
both the CUDA library and the CPU oracle consume.

Array shapes follow src/ModScbInit.f90:22-119 (Fortran order): ``*1`` arrays are
(nthe,npsi,nzeta+1), the others (nthe,npsi,nzeta).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

PI = 3.141592653589793238462643383279502884197


def _f(shape):
    return np.zeros(shape, dtype=np.float64, order="F")


def steffen_derivs(xa, ya, axis):
    """Derivative at the nodes of GSL's Steffen spline along ``axis`` (vectorised
    restatement of gsl interpolation/steffen.c as driven by src/RamGSL.c:255-277)."""
    y = np.moveaxis(np.asarray(ya, dtype=np.float64), axis, 0)
    n = y.shape[0]
    shp = (n,) + (1,) * (y.ndim - 1)
    h = np.diff(xa).reshape((n - 1,) + (1,) * (y.ndim - 1))
    s = np.diff(y, axis=0) / h
    yp = np.empty_like(y)
    yp[0] = s[0]
    sim1, si = s[:-1], s[1:]
    him1, hi = h[:-1], h[1:]
    p = (sim1 * hi + si * him1) / (him1 + hi)
    sgn = lambda v: np.where(v < 0, -1.0, 1.0)
    yp[1:-1] = (sgn(sim1) + sgn(si)) * np.minimum(np.abs(sim1), np.minimum(np.abs(si), 0.5 * np.abs(p)))
    yp[-1] = s[-1]
    d = yp.copy()
    hl, sl = h[-1], s[-1]
    a = (yp[-2] + yp[-1] - 2 * sl) / hl / hl
    b = (3 * sl - 2 * yp[-2] - yp[-1]) / hl
    delx = xa[-1] - xa[-2]
    d[-1] = yp[-2] + delx * (2.0 * b + delx * 3.0 * a)
    d[d == 0.0] = 1e-31
    return np.asfortranarray(np.moveaxis(d, 0, axis))


def derivs3d(thetaVal, rhoVal, zetaVal, f3):
    """GSL_Derivs 3-D driver (src/ModRamGSL.f90:794-869) on an (nthe,npsi,nzeta) array."""
    return (steffen_derivs(thetaVal, f3, 0), steffen_derivs(rhoVal, f3, 1), steffen_derivs(zetaVal, f3, 2))


@dataclass
class ScbInputs:
    nthe: int
    npsi: int
    nzeta: int
    isotropy: int
    thetaVal: np.ndarray
    rhoVal: np.ndarray
    zetaVal: np.ndarray
    psiVal: np.ndarray
    f: np.ndarray
    alphaVal: np.ndarray   # (nzeta+1)
    fzet: np.ndarray       # (nzeta+1)
    x: np.ndarray          # *1
    y: np.ndarray
    z: np.ndarray
    alfa: np.ndarray       # *1
    psi: np.ndarray        # *1
    pper: np.ndarray       # *1
    ppar: np.ndarray       # *1
    sigma: np.ndarray      # *1
    bsq0: np.ndarray       # *1 analytic dipole bsq used to build the pressure
    dPPerdTheta: np.ndarray = None
    dPPerdRho: np.ndarray = None
    dPPerdZeta: np.ndarray = None
    dBsqdTheta: np.ndarray = None
    dBsqdRho: np.ndarray = None
    dBsqdZeta: np.ndarray = None
    dPPerdPsi: np.ndarray = None
    dPPerdAlpha: np.ndarray = None
    dBsqdPsi: np.ndarray = None
    dBsqdAlpha: np.ndarray = None
    dPdAlpha: np.ndarray = None
    dPdPsi: np.ndarray = None
    chiVal: np.ndarray = None   # (nthe) field-line coordinate nodes, src/ModScbIO.f90:167
    extra: dict = field(default_factory=dict)


def build_scb(nthe=101, npsi=45, nzeta=97, constTheta=0.2, xpsiin=1.75, xpsiout=7.5, p0_nPa=2.0, aniso=0.5,
              isotropy=0, warp=0.0, seed=1) -> ScbInputs:
    """Dipole SCB state with an analytic anisotropic pressure.

    ``warp`` > 0 stretches the field lines tailward a little (a smooth, seeded
    non-dipolar perturbation of x,y,z) so the grid is non-orthogonal and every
    corner coefficient vec1/3/7/9 is exercised.
    """
    xzero = 6.6
    xzero3 = xzero ** 3
    x = _f((nthe, npsi, nzeta + 1))
    y = _f((nthe, npsi, nzeta + 1))
    z = _f((nthe, npsi, nzeta + 1))
    B = _f((nthe, npsi, nzeta + 1))
    # get_dipole_lines(radMin, radMax, constTheta, nthe, nR=npsi, nT=nzeta, ..., RAM=.false.)
    kk = np.arange(nthe)
    for i in range(npsi):
        r0 = xpsiin + i / (npsi - 1) * (xpsiout - xpsiin)
        t1 = np.arcsin(np.sqrt(1.0 / r0))
        t0 = PI - t1
        tt = t0 + kk / (nthe - 1) * (t1 - t0)
        tt = tt + constTheta * np.sin(2.0 * tt)
        rt = r0 * np.sin(tt) ** 2
        for j in range(1, nzeta):              # Fortran j = 2..nT
            zt = 2 * PI * (j - 1) / (nzeta - 1)
            x[:, i, j] = rt * np.cos(zt) * np.sin(tt)
            y[:, i, j] = rt * np.sin(zt) * np.sin(tt)
            z[:, i, j] = rt * np.cos(tt)
            B[:, i, j] = np.sqrt(1 + 3 * np.cos(tt) ** 2) / rt ** 3
    for a in (x, y, z, B):
        a[:, :, 0] = a[:, :, nzeta - 1]
        a[:, :, nzeta] = a[:, :, 1]
    if warp > 0.0:
        # tail-ward stretch growing with equatorial distance, strongest at midnight
        req = np.sqrt(x ** 2 + y ** 2 + z ** 2)
        phi = np.arctan2(y, x)
        s = warp * (req / xpsiout) ** 2 * (0.5 - 0.5 * np.cos(phi))
        x = np.asfortranarray(x - s * np.abs(x) * 0.3)
        z = np.asfortranarray(z * (1.0 - 0.5 * s))
        for a in (x, y, z):
            a[:, :, 0] = a[:, :, nzeta - 1]
            a[:, :, nzeta] = a[:, :, 1]

    thetaVal = PI * np.arange(nthe) / (nthe - 1)
    rhoVal = np.arange(npsi) / (npsi - 1)
    dphi = 2 * PI / (nzeta - 1)
    zetaVal = (np.arange(nzeta) - 1) * dphi
    # src/ModScbIO.f90:179-205
    xpsitot = xpsiout - xpsiin
    xpl = xpsiin + xpsitot * (np.arange(npsi) / (npsi - 1))
    psiVal = -xzero3 / xpl
    f = (xzero3 / xpl ** 2) * xpsitot
    alphaVal = (np.arange(nzeta + 1) - 1) * dphi
    fzet = np.ones(nzeta + 1)
    alfa = _f((nthe, npsi, nzeta + 1))
    psi = _f((nthe, npsi, nzeta + 1))
    alfa[:, :, :] = alphaVal[None, None, :]      # alfges
    psi[:, :, :] = psiVal[None, :, None]          # psiges

    # ---- pressure (normalised by pnormal, src/ModScbInit.f90:266) -----------------------
    bnormal = 0.31 / xzero3 * 1.0e5
    pnormal = bnormal * bnormal / (4.0 * PI * 1.0e-7) * 1.0e-9
    bf = xzero3 * B
    bsq0 = np.asfortranarray(bf ** 2)
    ieq = (nthe + 1) // 2 - 1                      # nThetaEquator (0-based)
    req = np.sqrt(x[ieq] ** 2 + y[ieq] ** 2)       # (npsi, nzeta+1)
    phi_eq = np.arctan2(y[ieq], x[ieq])
    pEq = (p0_nPa / pnormal) * (req / 4.0) ** -3.5 * np.exp(-((req - 4.0) ** 2) / 4.0) * (1.0 + 0.3 * np.cos(phi_eq))
    aratio = aniso * np.ones_like(pEq)
    ratioB = np.minimum(bf[ieq][None, :, :] / bf, 1.0)
    gParam = 1.0 / ((1.0 + aratio[None] * (1.0 - ratioB)) ** 2)     # iLossCone == 1, src/ModScbRun.f90:1107-1110
    if isotropy == 1:
        ppar = np.asfortranarray(np.broadcast_to(pEq[None], bf.shape).copy())
        pper = ppar.copy(order="F")
    else:
        ppar = np.asfortranarray(pEq[None] * 1.0 / (1.0 + 2.0 * aratio[None] / 3.0) * np.sqrt(gParam))
        pper = np.asfortranarray(pEq[None] * (aratio[None] + 1.0) / (1.0 + 2.0 * aratio[None] / 3.0) * gParam)
    sigma = np.asfortranarray(1.0 + (pper - ppar) / bsq0)

    inp = ScbInputs(nthe=nthe, npsi=npsi, nzeta=nzeta, isotropy=isotropy, thetaVal=thetaVal, rhoVal=rhoVal, zetaVal=zetaVal,
                    psiVal=psiVal, f=f, alphaVal=alphaVal, fzet=fzet, x=x, y=y, z=z, alfa=alfa, psi=psi, pper=pper, ppar=ppar,
                    sigma=sigma, bsq0=bsq0)
    inp.chiVal = thetaVal + constTheta * np.sin(2.0 * thetaVal)      # src/ModScbIO.f90:167
    # ---- derivatives the way `pressure` ends (src/ModScbRun.f90:1161-1175) ----------------
    p3 = pper[:, :, :nzeta]
    inp.dPPerdTheta, inp.dPPerdRho, inp.dPPerdZeta = derivs3d(thetaVal, rhoVal, zetaVal, p3)
    inp.dBsqdTheta, inp.dBsqdRho, inp.dBsqdZeta = derivs3d(thetaVal, rhoVal, zetaVal, bsq0[:, :, :nzeta])
    inp.dPPerdPsi = np.asfortranarray(inp.dPPerdRho / f[None, :, None])
    inp.dBsqdPsi = np.asfortranarray(inp.dBsqdRho / f[None, :, None])
    inp.dPPerdAlpha = np.asfortranarray(inp.dPPerdZeta / fzet[None, None, :nzeta])
    inp.dBsqdAlpha = np.asfortranarray(inp.dBsqdZeta / fzet[None, None, :nzeta])
    inp.dPdPsi = inp.dPPerdPsi.copy(order="F")      # isotropic branch inputs
    inp.dPdAlpha = inp.dPPerdAlpha.copy(order="F")
    inp.extra.update(bnormal=bnormal, pnormal=pnormal, xzero3=xzero3)
    return inp


def equatorial_pressure_fn(p0_nPa=2.0, aniso=0.5, xzero=6.6):
    """The 2-D front end of `pressure` (src/ModScbRun.f90:753-1086) for the synthetic workload: instead of
    interpolating RAM's PPerT/PParT to the equatorial foot points, evaluate SURVEY 8(d)'s analytic profile
    there.  Returns f(xEq, yEq) -> (pperEq, pparEq), normalised by pnormal, (npsi, nzeta+1), with the
    split of the equatorial pressure into p_perp / p_par used by build_scb."""
    bnormal = 0.31 / xzero ** 3 * 1.0e5
    pnormal = bnormal * bnormal / (4.0 * PI * 1.0e-7) * 1.0e-9

    def fn(xEq, yEq):
        req = np.sqrt(xEq ** 2 + yEq ** 2)
        phi = np.arctan2(yEq, xEq)
        pEq = (p0_nPa / pnormal) * (req / 4.0) ** -3.5 * np.exp(-((req - 4.0) ** 2) / 4.0) * (1.0 + 0.3 * np.cos(phi))
        ppar = pEq / (1.0 + 2.0 * aniso / 3.0)
        pper = pEq * (aniso + 1.0) / (1.0 + 2.0 * aniso / 3.0)
        for v in (pper, ppar):                       # periodic columns (k = 1 <- nzeta, nzeta+1 <- 2)
            v[:, 0] = v[:, -2]
            v[:, -1] = v[:, 1]
        return np.asfortranarray(pper), np.asfortranarray(ppar)

    return fn


def ram_field_lines(LZ, MLT, nthe=101, constTheta=0.2, wiggle=0.0, seed=3, outside_fraction=0.0):
    """Synthetic input of computehI's integral block (src/ModRamScb.f90:372-410): nR x nT dipole field lines through the
    RAM equatorial points (LZ(2:nR+1), MLT) sampled at the nthe nodes of chiVal = theta + constTheta sin(2 theta)
    (src/ModScbIO.f90:167), the field strength in units of bnormal and a hydrogen density along each line -- the role
    get_dipole_lines plays in the reference's 'DIPL' branch (:243-245).  wiggle > 0 adds a smooth perturbation of B with
    a few local extrema per line (the mirror search of src/RamGSL.c:545-566 is written for non-monotonic B) and moves
    the minimum off the equatorial node (the fix-up of :388-394); outside_fraction marks random lines outsideMGNP."""
    rng = np.random.default_rng(seed)
    LZ = np.asarray(LZ, dtype=np.float64)
    MLT = np.asarray(MLT, dtype=np.float64)
    nR, nT = len(LZ), len(MLT)
    theta = np.linspace(0.0, np.pi, nthe)
    chiVal = theta + constTheta * np.sin(2.0 * theta)
    shape = (nthe, nR, nT)
    x, y, z, b, dens = (np.zeros(shape, order="F") for _ in range(5))
    for i in range(nR):
        lam_foot = np.arccos(np.sqrt(1.0 / LZ[i]))
        # nodes at arc-length fraction chiVal / pi, as mapTheta leaves them (src/ModScbEuler.f90:15-75): computehI turns
        # the integrals over chi into integrals over arc length with the factor length / pi (src/ModRamScb.f90:402-405)
        lam_d = np.linspace(lam_foot, -lam_foot, 20001)
        ds = LZ[i] * np.cos(lam_d) * np.sqrt(1.0 + 3.0 * np.sin(lam_d) ** 2)          # |ds / dlambda| of a dipole line
        s_d = np.concatenate(([0.0], np.cumsum(0.5 * (ds[1:] + ds[:-1]) * np.abs(np.diff(lam_d)))))
        lam = np.interp(chiVal / np.pi * s_d[-1], s_d, lam_d)           # foot point ... equator ... conjugate foot point
        r = LZ[i] * np.cos(lam) ** 2
        bd = np.sqrt(1.0 + 3.0 * np.sin(lam) ** 2) / np.cos(lam) ** 6 / LZ[i] ** 3 * (30574.0 / 1.0)   # nT at the surface / bnormal = 1
        for j in range(nT):
            phi = MLT[j] * 2.0 * np.pi / 24.0 - np.pi
            x[:, i, j] = r * np.cos(lam) * np.cos(phi)
            y[:, i, j] = r * np.cos(lam) * np.sin(phi)
            z[:, i, j] = r * np.sin(lam)
            w = 1.0
            if wiggle:
                ph = rng.random(3) * 2.0 * np.pi
                w = 1.0 + wiggle * (np.sin(5.0 * theta + ph[0]) + 0.5 * np.sin(11.0 * theta + ph[1]) + 0.3 * np.sin(2.0 * theta + ph[2]))
            b[:, i, j] = bd * w
            dens[:, i, j] = 1.0e3 * (1.0 + 0.3 * rng.random()) * (LZ[i] / r) ** 3 * (1.0 + 0.1 * np.cos(3.0 * theta))
    outside = np.asfortranarray((rng.random((nR, nT)) < outside_fraction).astype(np.int32))
    return dict(chiVal=chiVal, xRAM=x, yRAM=y, zRAM=z, bRAM=b, density=dens, outsideMGNP=outside,
                nThetaEquator=nthe // 2 + 1, bnormal=1.0)


def synthetic_ram_pressures(g, seed=3):
    """Synthetic ring-current-like RAM pressures for the device front end of `pressure` (src/ModScbRun.f90:858-875):
    PPerT, PParT (nS,NR,NT) [keV/cm^3], the species%SCB flags, LZ(NR+1), PHI(NT).  Peak ~ 12 keV/cm^3 near L = 4 with a
    day-night asymmetry (the magnitudes of output/test1/pressure.ref), anisotropy p_par = 0.7 p_perp, 5 % seeded noise."""
    rng = np.random.default_rng(seed)
    LZ, PHI = g.LZ[:g.NR + 1], g.PHI[:g.NT]
    base = 12.0 * np.exp(-((LZ[:g.NR, None] - 4.0) / 1.2) ** 2) * (1 + 0.3 * np.cos(PHI[None, :]))
    frac = (1.0, 0.3, 0.1, 0.05)[:g.nS]
    PPerT = np.asfortranarray(np.stack([base * f * (1 + 0.05 * rng.random(base.shape)) for f in frac]))
    PParT = np.asfortranarray(0.7 * PPerT * (1 + 0.05 * rng.random(PPerT.shape)))
    scb = np.array([1, 1, 1, 0][:g.nS], dtype=np.int32)
    return PPerT, PParT, scb, LZ, PHI
