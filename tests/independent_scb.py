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
