"""A second, independent restatement (numpy, whole-array) of two SCB routines, written from the
Fortran text and not from oracle/scb_oracle.cpp: `metrica` (src/ModScbEquation.f90:18-280, the nine
stencil coefficients of the alpha equation from the 27-point neighbourhood of x, y, z) and one
lexicographic SOR sweep of `iterateAlpha` (src/ModScbEuler.f90:204-263).  The C++ oracle must agree
bit for bit (tests/test_cpu.py).  Arrays are [i, j, k] = (theta, psi, zeta), 0-based.
"""
import numpy as np


def spacing(nthe, npsi, nzeta):
    """src/ModScbInit.f90:131-148"""
    pi_d = 3.14159265358979323846264338327950288419716939937510
    dr = 1. / float(npsi - 1)
    dt = pi_d / float(nthe - 1)
    dp = 2 * pi_d / float(nzeta - 1)
    rdr, rdt, rdp = 1. / dr, 1. / dt, 1. / dp
    return dict(rdr=rdr, rdt=rdt, rdp=rdp, rdtsq=rdt ** 2, rdpsq=rdp ** 2, rdr2=0.5 * rdr, rdt2=0.5 * rdt, rdp2=0.5 * rdp,
                rdr4=0.25 * rdr, rdt4=0.25 * rdt, rdp4=0.25 * rdp, rdpdt4=0.25 * rdp * rdt)


def metrica(x, y, z, nthe, npsi, nzeta):
    """vecd, vec1..vec9 on the interior i=2..nthe-1, j=2..npsi-1, k=2..nzeta (zero elsewhere)."""
    s = spacing(nthe, npsi, nzeta)
    I = slice(1, nthe - 1)
    J = slice(1, npsi - 1)
    K = slice(1, nzeta)

    def sh(a, di, dj, dk):            # a(i+di, j+dj, k+dk) on the interior
        return a[1 + di:nthe - 1 + di, 1 + dj:npsi - 1 + dj, 1 + dk:nzeta + dk]

    def derivs(a):
        """(theta, zeta, rho) derivatives of one coordinate at the five points a..e (:88-152)."""
        t = {"a": (sh(a, 1, 0, 0) - sh(a, 0, 0, 0)) * s["rdt"],
             "b": (sh(a, 1, 0, 1) + sh(a, 1, 0, 0) - sh(a, -1, 0, 1) - sh(a, -1, 0, 0)) * s["rdt4"],
             "c": (sh(a, 0, 0, 0) - sh(a, -1, 0, 0)) * s["rdt"],
             "d": (sh(a, 1, 0, 0) + sh(a, 1, 0, -1) - sh(a, -1, 0, 0) - sh(a, -1, 0, -1)) * s["rdt4"],
             "e": (sh(a, 1, 0, 0) - sh(a, -1, 0, 0)) * s["rdt2"]}
        p = {"a": (sh(a, 1, 0, 1) + sh(a, 0, 0, 1) - sh(a, 1, 0, -1) - sh(a, 0, 0, -1)) * s["rdp4"],
             "b": (sh(a, 0, 0, 1) - sh(a, 0, 0, 0)) * s["rdp"],
             "c": (sh(a, 0, 0, 1) + sh(a, -1, 0, 1) - sh(a, 0, 0, -1) - sh(a, -1, 0, -1)) * s["rdp4"],
             "d": (sh(a, 0, 0, 0) - sh(a, 0, 0, -1)) * s["rdp"],
             "e": (sh(a, 0, 0, 1) - sh(a, 0, 0, -1)) * s["rdp2"]}
        r = {"a": (sh(a, 1, 1, 0) + sh(a, 0, 1, 0) - sh(a, 1, -1, 0) - sh(a, 0, -1, 0)) * s["rdr4"],
             "b": (sh(a, 0, 1, 1) + sh(a, 0, 1, 0) - sh(a, 0, -1, 1) - sh(a, 0, -1, 0)) * s["rdr4"],
             "c": (sh(a, 0, 1, 0) + sh(a, -1, 1, 0) - sh(a, 0, -1, 0) - sh(a, -1, -1, 0)) * s["rdr4"],
             "d": (sh(a, 0, 1, -1) + sh(a, 0, 1, 0) - sh(a, 0, -1, -1) - sh(a, 0, -1, 0)) * s["rdr4"],
             "e": (sh(a, 0, 1, 0) - sh(a, 0, -1, 0)) * s["rdr2"]}
        return t, p, r

    xt, xp, xr = derivs(x)
    yt, yp, yr = derivs(y)
    zt, zp, zr = derivs(z)
    aj, grs, gps, gts, grgp, gpgt, gtgr = {}, {}, {}, {}, {}, {}, {}
    for q in "abcde":
        aj[q] = (xr[q] * (yp[q] * zt[q] - yt[q] * zp[q]) + xp[q] * (yt[q] * zr[q] - yr[q] * zt[q])
                 + xt[q] * (yr[q] * zp[q] - yp[q] * zr[q]))
        grx = (yp[q] * zt[q] - yt[q] * zp[q]) / aj[q]
        gry = (zp[q] * xt[q] - zt[q] * xp[q]) / aj[q]
        grz = (xp[q] * yt[q] - xt[q] * yp[q]) / aj[q]
        gpx = (yt[q] * zr[q] - yr[q] * zt[q]) / aj[q]
        gpy = (zt[q] * xr[q] - zr[q] * xt[q]) / aj[q]
        gpz = (xt[q] * yr[q] - xr[q] * yt[q]) / aj[q]
        gtx = (yr[q] * zp[q] - yp[q] * zr[q]) / aj[q]
        gty = (zr[q] * xp[q] - zp[q] * xr[q]) / aj[q]
        gtz = (xr[q] * yp[q] - xp[q] * yr[q]) / aj[q]
        grs[q] = (grx ** 2 + gry ** 2 + grz ** 2)
        gps[q] = (gpx ** 2 + gpy ** 2 + gpz ** 2)
        gts[q] = (gtx ** 2 + gty ** 2 + gtz ** 2)
        grgp[q] = (gpx * grx + gpy * gry + gpz * grz)
        gpgt[q] = (gpx * gtx + gpy * gty + gpz * gtz)
        gtgr[q] = (gtx * grx + gty * gry + gtz * grz)
    v1 = {q: (grs[q] * gts[q] - gtgr[q] ** 2) * aj[q] * s["rdtsq"] for q in "abc"}
    v2 = {q: (grs[q] * gpgt[q] - grgp[q] * gtgr[q]) * aj[q] * s["rdpdt4"] for q in "abcd"}
    v3 = {q: (grs[q] * gps[q] - grgp[q] ** 2) * aj[q] * s["rdpsq"] for q in "bcd"}
    out = {n: np.zeros((nthe, npsi, nzeta)) for n in ("vecd", "vec1", "vec2", "vec3", "vec4", "vec6", "vec7", "vec8", "vec9")}
    out["vecd"][I, J, K] = (v1["a"] + v1["c"]) + (v3["b"] + v3["d"])
    out["vec1"][I, J, K] = (v2["c"] + v2["d"])
    out["vec2"][I, J, K] = (v2["c"] - v2["a"]) + v3["d"]
    out["vec3"][I, J, K] = -(v2["a"] + v2["d"])
    out["vec4"][I, J, K] = v1["c"] + (v2["d"] - v2["b"])
    out["vec6"][I, J, K] = v1["a"] + (v2["b"] - v2["d"])
    out["vec7"][I, J, K] = -(v2["c"] + v2["b"])
    out["vec8"][I, J, K] = v3["b"] + (v2["a"] - v2["c"])
    out["vec9"][I, J, K] = (v2["a"] + v2["b"])
    return out


def sor_alpha_sweeps(alfa, vec, vecx, nthe, npsi, nzeta, nT, nsweeps, om):
    """`nsweeps` lexicographic SOR sweeps of every psi surface (src/ModScbEuler.f90:213-244): k outer,
    iz inner, `om` = relaxation factor of each sweep (1 for the first).  Returns alfa and the max
    |resid| over iz = 2..nthe-1 (here: the updated columns) of the last sweep per surface."""
    a = alfa.copy()
    resmax = np.zeros(npsi)
    for jz in range(1, npsi - 1):
        for it in range(nsweeps):
            rm = 0.0
            for k in range(1, nzeta):
                for iz in range(nT, nthe - nT):
                    res = (-vec["vecd"][iz, jz, k] * a[iz, jz, k]
                           + vec["vec1"][iz, jz, k] * a[iz - 1, jz, k - 1] + vec["vec2"][iz, jz, k] * a[iz, jz, k - 1]
                           + vec["vec3"][iz, jz, k] * a[iz + 1, jz, k - 1] + vec["vec4"][iz, jz, k] * a[iz - 1, jz, k]
                           + vec["vec6"][iz, jz, k] * a[iz + 1, jz, k] + vec["vec7"][iz, jz, k] * a[iz - 1, jz, k + 1]
                           + vec["vec8"][iz, jz, k] * a[iz, jz, k + 1] + vec["vec9"][iz, jz, k] * a[iz + 1, jz, k + 1]
                           - vecx[iz, jz, k])
                    a[iz, jz, k] = a[iz, jz, k] + om[it] * (res / vec["vecd"][iz, jz, k])
                    rm = max(rm, abs(res))
            resmax[jz] = rm
    return a, resmax


def _point_metrics(xt, xp, xr, yt, yp, yr, zt, zp, zr):
    """Jacobian and the gradient products at the five points (:154-242 / :414-502, same text)."""
    aj, grs, gps, gts, grgp, gpgt, gtgr = {}, {}, {}, {}, {}, {}, {}
    for q in "abcde":
        aj[q] = (xr[q] * (yp[q] * zt[q] - yt[q] * zp[q]) + xp[q] * (yt[q] * zr[q] - yr[q] * zt[q])
                 + xt[q] * (yr[q] * zp[q] - yp[q] * zr[q]))
        grx = (yp[q] * zt[q] - yt[q] * zp[q]) / aj[q]
        gry = (zp[q] * xt[q] - zt[q] * xp[q]) / aj[q]
        grz = (xp[q] * yt[q] - xt[q] * yp[q]) / aj[q]
        gpx = (yt[q] * zr[q] - yr[q] * zt[q]) / aj[q]
        gpy = (zt[q] * xr[q] - zr[q] * xt[q]) / aj[q]
        gpz = (xt[q] * yr[q] - xr[q] * yt[q]) / aj[q]
        gtx = (yr[q] * zp[q] - yp[q] * zr[q]) / aj[q]
        gty = (zr[q] * xp[q] - zp[q] * xr[q]) / aj[q]
        gtz = (xr[q] * yp[q] - xp[q] * yr[q]) / aj[q]
        grs[q] = (grx ** 2 + gry ** 2 + grz ** 2)
        gps[q] = (gpx ** 2 + gpy ** 2 + gpz ** 2)
        gts[q] = (gtx ** 2 + gty ** 2 + gtz ** 2)
        grgp[q] = (gpx * grx + gpy * gry + gpz * grz)
        gpgt[q] = (gpx * gtx + gpy * gty + gpz * gtz)
        gtgr[q] = (gtx * grx + gty * gry + gtz * grz)
    return aj, grs, gps, gts, grgp, gpgt, gtgr


def metric(x, y, z, nthe, npsi, nzeta):
    """The psi-equation coefficients (src/ModScbEquation.f90:283-540): half points along theta (a, c)
    and along rho (b, d)."""
    s = spacing(nthe, npsi, nzeta)
    rdr = s["rdr"]
    rdrsq, rdtdr4 = rdr ** 2, 0.25 * s["rdt"] * rdr
    I, J, K = slice(1, nthe - 1), slice(1, npsi - 1), slice(1, nzeta)

    def sh(a, di, dj, dk):
        return a[1 + di:nthe - 1 + di, 1 + dj:npsi - 1 + dj, 1 + dk:nzeta + dk]

    def derivs(a):
        t = {"a": (sh(a, 1, 0, 0) - sh(a, 0, 0, 0)) * s["rdt"],
             "b": (sh(a, 1, 1, 0) + sh(a, 1, 0, 0) - sh(a, -1, 1, 0) - sh(a, -1, 0, 0)) * s["rdt4"],
             "c": (sh(a, 0, 0, 0) - sh(a, -1, 0, 0)) * s["rdt"],
             "d": (sh(a, 1, 0, 0) + sh(a, 1, -1, 0) - sh(a, -1, 0, 0) - sh(a, -1, -1, 0)) * s["rdt4"],
             "e": (sh(a, 1, 0, 0) - sh(a, -1, 0, 0)) * s["rdt2"]}
        p = {"a": (sh(a, 1, 0, 1) + sh(a, 0, 0, 1) - sh(a, 1, 0, -1) - sh(a, 0, 0, -1)) * s["rdp4"],
             "b": (sh(a, 0, 1, 1) + sh(a, 0, 0, 1) - sh(a, 0, 1, -1) - sh(a, 0, 0, -1)) * s["rdp4"],
             "c": (sh(a, 0, 0, 1) + sh(a, -1, 0, 1) - sh(a, 0, 0, -1) - sh(a, -1, 0, -1)) * s["rdp4"],
             "d": (sh(a, 0, 0, 1) + sh(a, 0, -1, 1) - sh(a, 0, 0, -1) - sh(a, 0, -1, -1)) * s["rdp4"],
             "e": (sh(a, 0, 0, 1) - sh(a, 0, 0, -1)) * s["rdp2"]}
        r = {"a": (sh(a, 1, 1, 0) + sh(a, 0, 1, 0) - sh(a, 1, -1, 0) - sh(a, 0, -1, 0)) * s["rdr4"],
             "b": (sh(a, 0, 1, 0) - sh(a, 0, 0, 0)) * rdr,
             "c": (sh(a, 0, 1, 0) + sh(a, -1, 1, 0) - sh(a, 0, -1, 0) - sh(a, -1, -1, 0)) * s["rdr4"],
             "d": (sh(a, 0, 0, 0) - sh(a, 0, -1, 0)) * rdr,
             "e": (sh(a, 0, 1, 0) - sh(a, 0, -1, 0)) * s["rdr2"]}
        return t, p, r

    xt, xp, xr = derivs(x)
    yt, yp, yr = derivs(y)
    zt, zp, zr = derivs(z)
    aj, grs, gps, gts, grgp, gpgt, gtgr = _point_metrics(xt, xp, xr, yt, yp, yr, zt, zp, zr)
    v1 = {q: (gpgt[q] ** 2 - gps[q] * gts[q]) * aj[q] * s["rdtsq"] for q in "abc"}
    v2 = {q: (grgp[q] * gpgt[q] - gps[q] * gtgr[q]) * aj[q] * rdtdr4 for q in "abcd"}
    v3 = {q: (grgp[q] ** 2 - grs[q] * gps[q]) * aj[q] * rdrsq for q in "bcd"}
    out = {n: np.zeros((nthe, npsi, nzeta)) for n in ("vecd", "vec1", "vec2", "vec3", "vec4", "vec6", "vec7", "vec8", "vec9")}
    out["vecd"][I, J, K] = (v1["a"] + v1["c"] + v3["b"] + v3["d"])
    out["vec1"][I, J, K] = (v2["c"] + v2["d"])
    out["vec2"][I, J, K] = ((v2["c"] - v2["a"]) + v3["d"])
    out["vec3"][I, J, K] = -(v2["a"] + v2["d"])
    out["vec4"][I, J, K] = (v1["c"] + (v2["d"] - v2["b"]))
    out["vec6"][I, J, K] = (v1["a"] + (v2["b"] - v2["d"]))
    out["vec7"][I, J, K] = -(v2["c"] + v2["b"])
    out["vec8"][I, J, K] = (v3["b"] + (v2["a"] - v2["c"]))
    out["vec9"][I, J, K] = (v2["b"] + v2["a"])
    return out


def bandjacob(inp, derivs3d):
    """computeBandJacob (src/ModScbCompute.f90:412-496) with a given GSL_Derivs restatement
    (ramscb_b200.scb_synthetic.derivs3d: numpy, independent of the C++ one)."""
    nthe, npsi, nzeta = inp.nthe, inp.npsi, inp.nzeta
    X = [derivs3d(inp.thetaVal, inp.rhoVal, inp.zetaVal, np.asfortranarray(a[:, :, :nzeta])) for a in (inp.x, inp.y, inp.z)]
    (dXT, dXR, dXZ), (dYT, dYR, dYZ), (dZT, dZR, dZZ) = X
    jac = dXR * (dYZ * dZT - dYT * dZZ) + dXZ * (dYT * dZR - dYR * dZT) + dXT * (dYR * dZZ - dYZ * dZR)
    gRX = (dYZ * dZT - dYT * dZZ) / jac
    gRY = (dZZ * dXT - dZT * dXZ) / jac
    gRZ = (dXZ * dYT - dXT * dYZ) / jac
    gZX = (dYT * dZR - dYR * dZT) / jac
    gZY = (dZT * dXR - dZR * dXT) / jac
    gZZ = (dXT * dYR - dXR * dYT) / jac
    gTX = (dYR * dZZ - dYZ * dZR) / jac
    gTY = (dZR * dXZ - dZZ * dXR) / jac
    gTZ = (dXR * dYZ - dXZ * dYR) / jac
    out = {"jacobian": jac,
           "GradRhoSq": gRX ** 2 + gRY ** 2 + gRZ ** 2,
           "GradRhoGradZeta": gRX * gZX + gRY * gZY + gRZ * gZZ,
           "GradRhoGradTheta": gRX * gTX + gRY * gTY + gRZ * gTZ,
           "GradThetaSq": gTX ** 2 + gTY ** 2 + gTZ ** 2,
           "GradThetaGradZeta": gTX * gZX + gTY * gZY + gTZ * gZZ,
           "GradZetaSq": gZX ** 2 + gZY ** 2 + gZZ ** 2}
    ff = inp.f[None, :, None] * inp.fzet[None, None, :nzeta]
    Bx, By, Bz = ff * dXT / jac, ff * dYT / jac, ff * dZT / jac
    bsq = (out["GradRhoSq"] * out["GradZetaSq"] - out["GradRhoGradZeta"] ** 2) * ff ** 2
    for a in (Bx, By, Bz, bsq):
        a[:, :, 0] = a[:, :, nzeta - 1]
    out.update(Bx=Bx, By=By, Bz=Bz, bsq=bsq)
    return out


def newk_aniso(inp, m, nthe, npsi, nzeta):
    """newk, anisotropic Picard branch (src/ModScbEquation.f90:578-590); m = bandjacob() output."""
    f2 = (inp.f ** 2)[None, :, None]
    fz = inp.fzet[None, None, :nzeta]
    sg = inp.sigma[:, :, :nzeta]
    xpz = (m["GradRhoSq"] * m["GradZetaSq"] - m["GradRhoGradZeta"] ** 2)
    xpt = (m["GradRhoSq"] * m["GradThetaGradZeta"] - m["GradRhoGradZeta"] * m["GradRhoGradTheta"])
    c0 = -(f2 * fz) / sg / m["bsq"]
    tz = inp.dPPerdZeta[:, :, :nzeta] + 0.5 * (1. - sg) * inp.dBsqdZeta[:, :, :nzeta]
    tt = inp.dPPerdTheta[:, :, :nzeta] + 0.5 * (1. - sg) * inp.dBsqdTheta[:, :, :nzeta]
    return m["jacobian"] / f2 * c0 * (tz * xpz + tt * xpt)


# ---------------------------------------------------------------------------------------------
# GSL_Interpolation_1D (Steffen) and the three re-gridding maps, restated a second time: whole-line
# numpy (searchsorted instead of the bisection loop, coefficient arrays a,b,c,d built for every
# interval as gsl's steffen_init does) straight from src/ModRamGSL.f90:240-311, src/RamGSL.c:111-174
# and the published steffen.c.
def interp1d_steffen(x1, f1, x2):
    x1 = np.asarray(x1, dtype=np.float64)
    f1 = np.asarray(f1, dtype=np.float64)
    keep = [0]
    for i in range(1, len(x1)):                   # monotonicity filter of the Fortran wrapper
        if x1[i] > x1[keep[-1]]:
            keep.append(i)
    xa, fa = x1[keep], f1[keep]
    n = len(xa)
    assert n >= 3
    h = np.diff(xa)
    s = np.diff(fa) / h
    yp = np.empty(n)
    yp[0] = s[0]
    p = (s[:-1] * h[1:] + s[1:] * h[:-1]) / (h[:-1] + h[1:])
    sgn = lambda v: np.where(v < 0, -1.0, 1.0)
    yp[1:-1] = (sgn(s[:-1]) + sgn(s[1:])) * np.minimum(np.abs(s[:-1]), np.minimum(np.abs(s[1:]), 0.5 * np.abs(p)))
    yp[-1] = s[-1]
    a = (yp[:-1] + yp[1:] - 2 * s) / h / h
    b = (3 * s - 2 * yp[:-1] - yp[1:]) / h
    x2 = np.asarray(x2, dtype=np.float64)
    out = np.empty_like(x2)
    lo = x2 <= xa[0]
    hi = x2 >= xa[-1]
    out[lo] = fa[0] + (x2[lo] - xa[0]) / (xa[1] - xa[0]) * (fa[1] - fa[0])
    out[hi] = fa[-1] + (x2[hi] - xa[-1]) / (xa[-2] - xa[-1]) * (fa[-2] - fa[-1])
    m = ~(lo | hi)
    idx = np.searchsorted(xa, x2[m], side="right") - 1      # xa[idx] <= x < xa[idx+1]
    dx = x2[m] - xa[idx]
    out[m] = fa[idx] + dx * (yp[idx] + dx * (b[idx] + dx * a[idx]))
    return out


def _wrap(a, nzeta):
    a[:, :, 0] = a[:, :, nzeta - 1]
    a[:, :, nzeta] = a[:, :, 1]


def map_alpha(x, y, z, alfa, alphaVal, nthe, npsi, nzeta):
    x, y, z, alfa = (np.array(a, order="F") for a in (x, y, z, alfa))
    for j in range(npsi):
        for i in range(nthe):
            ao = alfa[i, j, :].copy()
            for a in (x, y, z):
                a[i, j, 1:nzeta] = interp1d_steffen(ao, a[i, j, :].copy(), alphaVal[1:nzeta])
    for a in (x, y, z):
        _wrap(a, nzeta)
    alfa[:, :, :] = alphaVal[None, None, :]
    return x, y, z, alfa


def map_psi(x, y, z, psi, psiVal, nthe, npsi, nzeta):
    x, y, z, psi = (np.array(a, order="F") for a in (x, y, z, psi))
    for k in range(1, nzeta):
        for i in range(nthe):
            po = psi[i, :, k].copy()
            for a in (x, y, z):
                a[i, :, k] = interp1d_steffen(po, a[i, :, k].copy(), psiVal)
    for a in (x, y, z):
        _wrap(a, nzeta)
    psi[:, :, :] = psiVal[None, :, None]
    return x, y, z, psi


def map_theta(x, y, z, chiVal, nthe, npsi, nzeta):
    x, y, z = (np.array(a, order="F") for a in (x, y, z))
    PI = 3.141592653589793238462643383279502884197
    for k in range(1, nzeta):
        for j in range(npsi):
            xo, yo, zo = x[:, j, k].copy(), y[:, j, k].copy(), z[:, j, k].copy()
            seg = np.sqrt((xo[1:] - xo[:-1]) ** 2 + (yo[1:] - yo[:-1]) ** 2 + (zo[1:] - zo[:-1]) ** 2)
            dist = np.zeros(nthe)
            for i in range(1, nthe):              # running sum in the reference's order
                dist[i] = dist[i - 1] + seg[i - 1]
            chiOld = dist / dist[-1] * PI
            for a, old in ((x, xo), (y, yo), (z, zo)):
                a[:, j, k] = interp1d_steffen(chiOld, old, chiVal)
    for a in (x, y, z):
        _wrap(a, nzeta)
    return x, y, z


# ---------------------------------------------------------------------------------------------
# `pressure`, anisotropic mapping from the equatorial pressures (src/ModScbRun.f90:1087-1160), whole-array
def pressure_aniso(pperEq, pparEq, bf, bsq, nthe, npsi, nzeta, iLossCone=1, iReduce=0):
    ieq = (nthe + 1) // 2 - 1
    pe = pperEq[None, :, :nzeta]
    pa = pparEq[None, :, :nzeta]
    bfk, bsqk = bf[:, :, :nzeta], bsq[:, :, :nzeta]

    def point(pe, pa, clamp, pEq_keep=None):
        pEq = (2.0 * pe + pa) / 3.0 if pEq_keep is None else pEq_keep
        ar = pe / pa - 1.0
        aL = -ar / (ar + 1)
        rB = bfk[ieq][None] / bfk
        if clamp:
            rB = np.where(rB < 1.0, rB, 1.0)
        if iLossCone == 2:
            q = bfk[0][None] / bfk
            rBI = np.where(q > 1.0 + 1.0e-9, q, 1.0 + 1.0e-9)
            frac = (rB + aL * rB) / (rBI + aL * rB)
            pparN = pa * (1.0 - frac)
            pperN = pe * (1.0 - frac)
            aN = pparN / pperN - 1.0
            ppar = pparN * (aN + 1.0) / (1.0 + aN * rB) * np.sqrt((rBI - 1.0) / (rBI - rB)) * (1.0 - (1.0 + aN * rB) / (rBI + aN * rB))
            pper = ppar / (1.0 + aN * rB)
        else:
            t = 1.0 + ar * (1.0 - rB)
            g = 1.0 / (t * t)
            ppar = pEq * 1.0 / (1.0 + 2.0 * ar / 3.0) * np.sqrt(g)
            pper = pEq * (ar + 1.0) / (1.0 + 2.0 * ar / 3.0) * g
        sigma = 1.0 + (pper - ppar) / bsqk
        tau = 1.0 - 2.0 * (pper - ppar) / bsqk * pper / ppar
        return pper, ppar, sigma, tau

    out = point(pe, pa, True)
    if iReduce == 1:
        unstable = out[3][ieq] < 0.0                      # (npsi, nzeta)
        pEq = (2.0 * pe + pa) / 3.0
        bE = bsqk[ieq][None]
        sixth = float(np.float32(1.0) / np.float32(6.0))  # the reference's single-precision 1./6.
        pe2 = sixth * (3.0 * pEq - bE + np.sqrt(bE * bE + 12.0 * bE * pEq + 9.0 * (pEq * pEq)))
        pa2 = 3.0 * pEq - 2.0 * pe2
        red = point(pe2, pa2, False, pEq_keep=pEq if iLossCone == 1 else None)
        out = tuple(np.where(unstable[None], r, o) for r, o in zip(red, out))
    return out
