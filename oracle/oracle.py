"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE -- see the header of
oracle/ram_oracle.cpp).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")

F_WPI, F_COULOMB, F_EMIC = 1, 2, 4


def build(force: bool = False) -> None:
    """Compile the oracle libraries (g++, a few seconds)."""
    cmd = ["make", "-C", _HERE, "-s"] + (["-B"] if force else [])
    subprocess.run(cmd, check=True)


def _load(name):
    path = os.path.join(_BUILD, name)
    if not os.path.exists(path):
        build()
    return C.CDLL(path)


_ram = None
_ram_variants = {}


def ram_lib(variant=""):
    """variant "fma": the same source built with FMA contraction allowed (oracle/Makefile) -- a measuring stick for the
    reference's own compiler-dependent rounding, never the parity oracle."""
    global _ram
    if variant:
        if variant not in _ram_variants:
            _ram_variants[variant] = _ram_lib_from("libram_oracle_" + variant + ".so")
        return _ram_variants[variant]
    if _ram is None:
        _ram = _ram_lib_from("libram_oracle.so")
    return _ram


def _ram_lib_from(name):
    if True:
        lib = _load(name)
        lib.orc_create.restype = C.c_void_p
        lib.orc_create.argtypes = [C.c_int] * 5
        lib.orc_destroy.argtypes = [C.c_void_p]
        lib.orc_set_array.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        lib.orc_set_iarray.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        lib.orc_set_scalar.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        for f in ("driftpara", "driftr", "driftp", "drifte", "driftmu", "cepara", "charexchange", "atmol",
                  "wavelo", "coulpara", "coulen", "coulmu", "sumrc", "anisch", "para_flc"):
            fn = getattr(lib, "orc_" + f)
            fn.argtypes = [C.c_void_p, C.c_int]
            fn.restype = None
        lib.orc_wpadif.argtypes = [C.c_void_p, C.c_int]
        lib.orc_wpadif.restype = C.c_long
        lib.orc_flcscatter.argtypes = [C.c_void_p, C.c_int]
        lib.orc_flcscatter.restype = C.c_long
        lib.orc_geosb.argtypes = [C.c_void_p, C.c_int]
        lib.orc_geosb.restype = None
        lib.orc_get_electric_field.argtypes = [C.c_void_p, C.c_int]
        lib.orc_get_electric_field.restype = None
        lib.orc_anisch_diffcoef.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.orc_anisch_diffcoef.restype = C.c_int
        lib.orc_ram_run.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.orc_ram_run.restype = C.c_double
        lib.orc_get_cdrift.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        for f in ("gcoul", "funt", "funi"):
            fn = getattr(lib, "orc_" + f)
            fn.argtypes = [C.c_double]
            fn.restype = C.c_double
        lib.orc_max_threads.restype = C.c_int
    return lib


def _f(shape, dtype=np.float64):
    return np.zeros(shape, dtype=dtype, order="F")


class RamOracle:
    """Holds one complete RAM state (reference layouts) and runs the restated
    operators on it.  Species index S is 1-based, as in the reference."""

    def __init__(self, g, inp, DTs=5.0, variant=""):
        self.lib = ram_lib(variant)
        self.g = g
        nS, NR, NT, NE, NPA = g.nS, g.NR, g.NT, g.NE, g.NPA
        self.h = self.lib.orc_create(nS, NR, NT, NE, NPA)
        self.arr = {}
        # grids
        for name in ("LZ", "RLZ", "EKEV", "WE", "DE", "EBND", "MU", "WMU", "DMU", "UPA", "RMAS",
                     "GREL", "GRBND", "V", "VBND", "EPP", "ERNH", "FFACTOR"):
            self._set(name, np.asfortranarray(getattr(g, name), dtype=np.float64).copy(order="F"))
        self._seti("QS", g.QS.copy())
        self._seti("kind", g.kind.copy())
        self._seti("khi", g.khi.copy())
        # fields
        for name in ("BNES", "dBdt", "VT", "EIR", "EIP", "FNHS", "FNIS", "BOUNHS", "BOUNIS", "HDNS", "dIdt",
                     "dIbndt", "NECR", "FGEOS", "F2", "WALOS1", "WALOS2", "WALOS3"):
            self._set(name, np.asfortranarray(getattr(inp, name), dtype=np.float64).copy(order="F"))
        self._seti("outsideMGNP", np.asfortranarray(inp.outsideMGNP, dtype=np.int32).copy(order="F"))
        # work / outputs
        self._set("CHARGE", _f((nS, NR, NT, NE, NPA)))
        self._set("FLUX", _f((nS, NR, NT, NE, NPA)))
        self._set("ATLOS", _f((nS, NR, NE)))
        for name in ("ATAW", "ATAC", "ATAW_emic_h", "ATAW_emic_he", "FLC_coef"):
            self._set(name, _f((NR, NT, NE, NPA)))
        for name in ("r_curvEq", "zeta1Eq", "zeta2Eq"):          # outputs of FLC_Radius, inputs of PARA_FLC
            self._set(name, _f((NR, NT)))
        for name in ("COULE", "COULI", "ATA", "GTA", "CEDR", "CIDR"):
            self._set(name, _f((nS, NE, NPA)))
        for name in ("DtDriftR", "DtDriftP", "DtDriftE", "DtDriftMu", "SETRC", "ELORC", "LSDR", "LSCHA", "LSATM",
                     "LSWAE", "LSCOE", "LSCSC"):
            self._set(name, _f((nS,)))
        self._set("PPERT", _f((nS, NR, NT)))
        self._set("PPART", _f((nS, NR, NT)))
        for name, v in (("MDR", g.MDR), ("DPHI", g.DPHI), ("CONF1", g.CONF1), ("CONF2", g.CONF2),
                        ("Kp", inp.Kp), ("Kpmax12", inp.Kpmax12), ("DTs", DTs), ("T", 0.0), ("Dt_bc", 300.0)):
            self.set_scalar(name, v)

    def _set(self, name, a):
        assert a.flags.f_contiguous and a.dtype == np.float64
        self.arr[name] = a
        self.lib.orc_set_array(self.h, name.encode(), a.ctypes.data)

    def _seti(self, name, a):
        assert a.dtype == np.int32
        self.arr[name] = a
        self.lib.orc_set_iarray(self.h, name.encode(), a.ctypes.data)

    def set_scalar(self, name, v):
        self.lib.orc_set_scalar(self.h, name.encode(), float(v))

    def set_array(self, name, a):
        """Replace the contents of a registered array."""
        self.arr[name][...] = a

    def __getattr__(self, name):
        arr = self.__dict__.get("arr", {})
        if name in arr:
            return arr[name]
        raise AttributeError(name)

    def op(self, name, S):
        return getattr(self.lib, "orc_" + name)(self.h, int(S))

    def cdrift(self, S, which):
        g = self.g
        out = _f((g.NR, g.NT, g.NE, g.NPA))
        self.lib.orc_get_cdrift(self.h, S, which, out.ctypes.data)
        return out

    def geosb(self, S, FluxLanl, s_comp):
        """GEOSB (src/ModRamBoundary.f90:241-319, boundary LANL): FGEOS(S,:,:,:) from the geosynchronous flux (NT,NE)"""
        self._set("FluxLanl", np.asfortranarray(FluxLanl, dtype=np.float64).copy(order="F"))
        self.set_scalar("s_comp", s_comp)
        self.lib.orc_geosb(self.h, int(S))

    def get_electric_field(self, vols, VTOL=None, VTN=None, t=0.0, TOLV=0.0, DtEfi=1.0, PHI=None, PHIOFS=0.0):
        """get_electric_field (src/ModRamEField.f90:14-63) -> VT"""
        if vols:
            self._set("PHI", np.asfortranarray(PHI, dtype=np.float64).copy())
            self.set_scalar("PHIOFS", PHIOFS)
        else:
            self._set("VTOL", np.asfortranarray(VTOL, dtype=np.float64).copy(order="F"))
            self._set("VTN", np.asfortranarray(VTN, dtype=np.float64).copy(order="F"))
            for n, v in (("TimeRamElapsed", t), ("TOLV", TOLV), ("DtEfi", DtEfi)):
                self.set_scalar(n, v)
        self.lib.orc_get_electric_field(self.h, 1 if vols else 0)

    def anisch_diffcoef(self, S, flags, t, AE=0, use_bas=True):
        """second half of ANISCH (src/ModRamRun.f90:422-605) from the table dict `t` (synthetic.synthetic_wave_tables);
        fills ATAW / ATAC or ATAW_emic_h / ATAW_emic_he; returns the GSLerr count"""
        for name in ("ENOR", "fpofc", "NDAAJ", "CDAAR", "BDAAR", "EKEV_emic", "fp2c_emic", "Daa_emic_h", "Daa_emic_he", "Ihs_emic",
                     "Ihes_emic", "XNE", "PAbn"):
            self._set(name, np.asfortranarray(t[name], dtype=np.float64).copy(order="F"))
        for name in ("ENG", "NCF", "ENG_emic", "NCF_emic"):
            self.set_scalar(name, t[name])
        self.set_scalar("AE", AE)
        self.set_scalar("DoUseBASdiff", 1.0 if use_bas else 0.0)
        self.set_scalar("electron_species", 1 + int(np.argmax(self.g.kind == 3)))
        return self.lib.orc_anisch_diffcoef(self.h, int(S), int(flags))

    def ram_run(self, flags=0, nthreads=None):
        if nthreads is None:
            nthreads = min(self.g.nS, self.lib.orc_max_threads())
        return self.lib.orc_ram_run(self.h, flags, nthreads)

    def __del__(self):
        try:
            self.lib.orc_destroy(self.h)
        except Exception:
            pass


# =============================================================================
# SCB oracle (oracle/scb_oracle.cpp)
# =============================================================================
_scb = None

SCB_OUT3 = ("derivXTheta", "derivXRho", "derivXZeta", "derivYTheta", "derivYRho", "derivYZeta", "derivZTheta", "derivZRho",
            "derivZZeta", "jacobian", "gradRhoX", "gradRhoY", "gradRhoZ", "gradZetaX", "gradZetaY", "gradZetaZ", "gradThetaX",
            "gradThetaY", "gradThetaZ", "GradRhoSq", "GradThetaSq", "GradZetaSq", "GradRhoGradTheta", "GradRhoGradZeta",
            "GradThetaGradZeta", "Bx", "By", "Bz", "vecd", "vec1", "vec2", "vec3", "vec4", "vec6", "vec7", "vec8", "vec9", "vecx",
            "vecr", "jGradRho", "jGradZeta", "jGradTheta", "Jx", "Jy", "Jz", "GradPx", "GradPy", "GradPz", "jCrossB", "GradP")
SCB_IN3 = ("dPPerdTheta", "dPPerdRho", "dPPerdZeta", "dBsqdTheta", "dBsqdRho", "dBsqdZeta", "dPPerdPsi", "dPPerdAlpha",
           "dBsqdPsi", "dBsqdAlpha", "dPdAlpha", "dPdPsi")
SCB_IN3P = ("x", "y", "z", "alfa", "psi", "pper", "ppar", "sigma")


def scb_lib():
    global _scb
    if _scb is None:
        lib = _load("libscb_oracle.so")
        lib.scbo_create.restype = C.c_void_p
        lib.scbo_create.argtypes = [C.c_int] * 3
        lib.scbo_destroy.argtypes = [C.c_void_p]
        lib.scbo_set_array.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        lib.scbo_set_scalar.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        lib.scbo_set_int.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        lib.scbo_get_scalar.argtypes = [C.c_void_p, C.c_char_p]
        lib.scbo_get_scalar.restype = C.c_double
        lib.scbo_interp1d.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.scbo_interp1d.restype = C.c_int
        for f in ("bandjacob", "convergence", "map_alpha", "map_psi", "map_theta"):
            getattr(lib, "scbo_" + f).argtypes = [C.c_void_p]
            getattr(lib, "scbo_" + f).restype = C.c_int
        for f in ("metrica", "metric", "newk", "newj", "pressure_aniso"):
            getattr(lib, "scbo_" + f).argtypes = [C.c_void_p]
            getattr(lib, "scbo_" + f).restype = None
        for f in ("iterate_alpha", "iterate_psi"):
            getattr(lib, "scbo_" + f).argtypes = [C.c_void_p, C.c_void_p]
            getattr(lib, "scbo_" + f).restype = C.c_int
        lib.scbo_pressure_raw.argtypes = [C.c_int] * 3 + [C.c_void_p] * 5 + [C.c_int] * 3 + [C.c_void_p] * 4
        lib.scbo_pressure_raw.restype = C.c_int
        lib.scbo_pressure_eq.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 4 + [C.c_double,
                                                                                                                            C.c_void_p, C.c_void_p]
        lib.scbo_pressure_eq.restype = None
        lib.scbo_derivs3d.argtypes = [C.c_void_p] * 5
        lib.scbo_steffen.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _scb = lib
    return _scb


class ScbOracle:
    """SCB state in the reference's layouts + the restated routines."""

    def __init__(self, inp):
        self.lib = scb_lib()
        self.inp = inp
        nthe, npsi, nzeta = inp.nthe, inp.npsi, inp.nzeta
        self.h = self.lib.scbo_create(nthe, npsi, nzeta)
        self.arr = {}
        for n in ("thetaVal", "rhoVal", "zetaVal", "psiVal", "f", "alphaVal", "fzet", "chiVal"):
            if getattr(inp, n, None) is None:
                continue
            self._set(n, np.ascontiguousarray(getattr(inp, n), dtype=np.float64).copy())
        for n in SCB_IN3P:
            self._set(n, np.asfortranarray(getattr(inp, n)).copy(order="F"))
        for n in SCB_IN3:
            self._set(n, np.asfortranarray(getattr(inp, n)).copy(order="F"))
        for n in ("bsq", "bf"):
            self._set(n, _f((nthe, npsi, nzeta + 1)))
        for n in SCB_OUT3:
            self._set(n, _f((nthe, npsi, nzeta)))
        self.set_int("isotropy", inp.isotropy)

    def _set(self, name, a):
        self.arr[name] = a
        self.lib.scbo_set_array(self.h, name.encode(), a.ctypes.data)

    def set_scalar(self, name, v):
        self.lib.scbo_set_scalar(self.h, name.encode(), float(v))

    def set_int(self, name, v):
        self.lib.scbo_set_int(self.h, name.encode(), int(v))

    def get(self, name):
        return self.lib.scbo_get_scalar(self.h, name.encode())

    def __getattr__(self, name):
        arr = self.__dict__.get("arr", {})
        if name in arr:
            return arr[name]
        raise AttributeError(name)

    def bandjacob(self):
        return self.lib.scbo_bandjacob(self.h)

    def metrica(self): self.lib.scbo_metrica(self.h)
    def metric(self): self.lib.scbo_metric(self.h)
    def newk(self): self.lib.scbo_newk(self.h)
    def newj(self): self.lib.scbo_newj(self.h)

    def iterate_alpha(self):
        ni = np.zeros(self.inp.npsi, dtype=np.int32)
        fail = self.lib.scbo_iterate_alpha(self.h, ni.ctypes.data)
        return fail, ni

    def iterate_psi(self):
        ni = np.zeros(self.inp.nzeta, dtype=np.int32)
        fail = self.lib.scbo_iterate_psi(self.h, ni.ctypes.data)
        return fail, ni

    def convergence(self):
        return self.lib.scbo_convergence(self.h)

    def pressure_aniso(self, pperEq, pparEq, iLossCone=1, iReduceAnisotropy=0):
        """Tail of `pressure` (src/ModScbRun.f90:1087-1175) from normalised equatorial pressures (npsi, nzeta+1)."""
        for n, a in (("pperEq", pperEq), ("pparEq", pparEq)):
            self._set(n, np.asfortranarray(a, dtype=np.float64).copy(order="F"))
        if "tau" not in self.arr:
            self._set("tau", _f((self.inp.nthe, self.inp.npsi, self.inp.nzeta + 1)))
        self.set_int("iLossCone", iLossCone)
        self.set_int("iReduceAnisotropy", iReduceAnisotropy)
        self.lib.scbo_pressure_aniso(self.h)

    PRESS_MODES = {"SKD": 0, "ROE": 1, "EXT": 2, "FLT": 3}

    def pressure_raw(self, PPerT, PParT, scb, LZ, PHI, PressMode="SKD", iSm2=4, SavGolIters=11):
        """front end of `pressure`, part 1 (src/ModScbRun.f90:858-980): RAM pressures summed over the species%SCB species,
        extended radially and smoothed; kept for pressure_front()."""
        a, b = np.asfortranarray(PPerT, dtype=np.float64), np.asfortranarray(PParT, dtype=np.float64)
        nS, NR, NT = a.shape
        nX = NR + 2 * int(np.floor(1.5 / (5. / NR)))
        flags = np.ascontiguousarray(scb, dtype=np.int32)
        lz, ph = np.ascontiguousarray(LZ, dtype=np.float64), np.ascontiguousarray(PHI, dtype=np.float64)
        r2, az = np.zeros(nX), np.zeros(NT)
        per, par = _f((nX, NT)), _f((nX, NT))
        rc = self.lib.scbo_pressure_raw(nS, NR, NT, a.ctypes.data, b.ctypes.data, flags.ctypes.data, lz.ctypes.data, ph.ctypes.data,
                                        self.PRESS_MODES[PressMode], iSm2, SavGolIters, r2.ctypes.data, az.ctypes.data,
                                        per.ctypes.data, par.ctypes.data)
        assert rc == 0, "unsupported PressMode / iSm2"
        self._raw = (r2, az, per, par)
        return self._raw

    def pressure_front(self):
        """front end of `pressure`, part 2 (:838-850, :1060-1086): normalised equatorial pressures (npsi, nzeta+1)"""
        nthe, npsi, nzeta = self.inp.nthe, self.inp.npsi, self.inp.nzeta
        ieq = (nthe + 1) // 2 - 1
        xe, ye = np.array(self.x[ieq], order="F"), np.array(self.y[ieq], order="F")
        r2, az, per, par = self._raw
        pe, pa = _f((npsi, nzeta + 1)), _f((npsi, nzeta + 1))
        self.lib.scbo_pressure_eq(npsi, nzeta, xe.ctypes.data, ye.ctypes.data, len(r2), len(az), r2.ctypes.data, az.ctypes.data,
                                  per.ctypes.data, par.ctypes.data, self.get("pnormal"), pe.ctypes.data, pa.ctypes.data)
        return pe, pa

    def map_alpha(self): return self.lib.scbo_map_alpha(self.h)
    def map_psi(self): return self.lib.scbo_map_psi(self.h)
    def map_theta(self): return self.lib.scbo_map_theta(self.h)

    def scb_run(self, pressure_fn, InConAlpha=1e-6, InConPsi=1e-6, blendInitial=0.5, blendMin=0.01, blendMax=1.0, damp=0.9,
                decreaseConvAlpha=0.5, decreaseConvPsi=0.5, nimax=5001, numit=200, MinSCBIterations=11, iLossCone=1,
                iReduceAnisotropy=0):
        """The outer iteration of scb_run (src/ModScbRun.f90:134-440; method = 2, iAMR = 0) composed from the
        restated routines.  pressure_fn(xEq, yEq) -> (pperEq, pparEq) is the 2-D front end of `pressure`."""
        nthe, npsi, nzeta = self.inp.nthe, self.inp.npsi, self.inp.nzeta
        ieq = (nthe + 1) // 2 - 1
        self.set_scalar("InConAlpha", InConAlpha); self.set_scalar("InConPsi", InConPsi); self.set_int("nimax", nimax)

        def pressure():
            if pressure_fn is None:          # the restated front end (pressure_raw before the call)
                pe, pa = self.pressure_front()
            else:
                pe, pa = pressure_fn(np.array(self.x[ieq], order="F"), np.array(self.y[ieq], order="F"))
            self.pressure_aniso(pe, pa, iLossCone, iReduceAnisotropy)

        def minjac():
            j = self.jacobian[1:nthe - 1, 1:npsi - 1, 1:nzeta]
            return -1e300 if np.isnan(j).any() else float(j.min())

        out = {"blendRetries": 0}
        start = {n: getattr(self, n).copy() for n in ("x", "y", "z", "alfa", "psi")}           # :134-143
        fail = self.bandjacob() != 0
        pressure()
        fail |= self.convergence() != 0
        out["normStart"] = (self.get("normDiff"), self.get("normJxB"), self.get("normGradP"))
        blendAlpha = blendPsi = blendInitial                                                   # :178-180
        errorAlpha = errorPsi = 0.0
        iteration, iConvGlobal = 1, 0
        alfaSav1, psiSav1 = self.alfa.copy(), self.psi.copy()                                   # :206-207
        while not fail:
            if self.bandjacob() != 0:                                                           # equation 1, :213-262
                fail = True; break
            pressure(); self.metrica(); self.newk()
            blendAlpha = min(max(blendAlpha, blendMin), blendMax)
            f, _ = self.iterate_alpha()
            if f:
                fail = True; break
            out["nisaveAlpha"], out["sumdbAlpha"] = int(self.get("nisave")), self.get("sumdb")
            errorAlpha = self.get("diffmx")
            prev = {n: getattr(self, n).copy() for n in ("x", "y", "z")}
            alphaPrev = self.alfa.copy()
            while True:
                self.alfa[...] = alphaPrev * blendAlpha + alfaSav1 * (1.0 - blendAlpha)
                if self.map_alpha() or self.map_theta():
                    fail = True; break
                fail |= self.bandjacob() != 0
                if minjac() < 0.0:
                    for n in ("x", "y", "z"):
                        getattr(self, n)[...] = prev[n]
                    blendAlpha = damp * blendAlpha
                    out["blendRetries"] += 1
                    if blendAlpha < blendMin:
                        fail = True; break
                    continue
                break
            if fail:
                break
            if self.bandjacob() != 0:                                                           # equation 2, :285-349
                fail = True; break
            pressure()
            fail |= self.convergence() != 0
            self.metric(); self.newj()
            blendPsi = min(max(blendPsi, blendMin), blendMax)
            f, _ = self.iterate_psi()
            if f:
                fail = True; break
            out["nisavePsi"], out["sumdbPsi"] = int(self.get("nisave")), self.get("sumdb")
            errorPsi = self.get("diffmx")
            prev = {n: getattr(self, n).copy() for n in ("x", "y", "z")}
            psiPrev = self.psi.copy()
            while True:
                self.psi[...] = psiPrev * blendPsi + (1.0 - blendPsi) * psiSav1
                if self.map_psi() or self.map_theta():
                    fail = True; break
                fail |= self.bandjacob() != 0
                if minjac() < 0.0:
                    for n in ("x", "y", "z"):
                        getattr(self, n)[...] = prev[n]
                    blendPsi = damp * blendPsi
                    out["blendRetries"] += 1
                    if blendPsi < blendMin:
                        fail = True; break
                    continue
                break
            if fail:
                break
            if errorAlpha < decreaseConvAlpha and errorPsi < decreaseConvPsi:                   # :363-378
                iConvGlobal = 1
            if (iteration < numit and iConvGlobal == 0) or iteration < MinSCBIterations:
                iteration += 1
                continue
            break
        out.update(iterations=iteration, iConvGlobal=iConvGlobal, SORFail=int(fail), blendAlpha=blendAlpha, blendPsi=blendPsi,
                   errorAlpha=errorAlpha, errorPsi=errorPsi)
        if fail:                                                                                # :397-413
            for n, v in start.items():
                getattr(self, n)[...] = v
            return out
        self.bandjacob(); pressure(); self.convergence()                                        # :427-429
        out["norm"] = (self.get("normDiff"), self.get("normJxB"), self.get("normGradP"))
        return out

    def derivs3d(self, f3):
        f3 = np.asfortranarray(f3, dtype=np.float64)
        out = [_f(f3.shape) for _ in range(3)]
        self.lib.scbo_derivs3d(self.h, f3.ctypes.data, *[o.ctypes.data for o in out])
        return out

    def __del__(self):
        try:
            self.lib.scbo_destroy(self.h)
        except Exception:
            pass


# ---- computehI integral block (oracle/hi_oracle.cpp) ---------------------------------------------------------
_hi = None


def hi_lib():
    global _hi
    if _hi is None:
        _hi = _load("libhi_oracle.so")
        _hi.hio_line.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 7
        _hi.hio_integrals.argtypes = [C.c_int] * 5 + [C.c_double] + [C.c_void_p] * 13
    return _hi


def hi_line(mirror, cVal, bf, var):
    """integrator_c + bounceaverage_c on one line; returns (mirror after the calls, yI, yH, yV)."""
    m = np.array(mirror, dtype=np.float64)
    c, b, v = (np.ascontiguousarray(a, dtype=np.float64) for a in (cVal, bf, var))
    yI, yH, yV = (np.zeros(len(m)) for _ in range(3))
    hi_lib().hio_line(len(c), len(m), m.ctypes.data, c.ctypes.data, b.ctypes.data, v.ctypes.data, yI.ctypes.data, yH.ctypes.data,
                      yV.ctypes.data)
    return m, yI, yH, yV


def hi_integrals(chiVal, mu, xRAM, yRAM, zRAM, bRAM, density, outsideMGNP, nThetaEquator, bnormal, HDens_cart=None):
    """src/ModRamScb.f90:372-410; returns I_cart, H_cart, HDens_cart, bZEq_Cart, bfMirror."""
    nthe, nR, nT = bRAM.shape
    nPa = len(mu)
    f = lambda a: np.asfortranarray(a, dtype=np.float64)
    chiVal, mu, xRAM, yRAM, zRAM, bRAM, density = (f(a) for a in (chiVal, mu, xRAM, yRAM, zRAM, bRAM, density))
    out = np.asfortranarray(outsideMGNP, dtype=np.int32)
    I, H, M = (_f((nR, nT, nPa)) for _ in range(3))
    D = _f((nR, nT, nPa)) if HDens_cart is None else f(HDens_cart).copy(order="F")
    bz = _f((nR, nT))
    hi_lib().hio_integrals(nthe, nR, nT, nPa, int(nThetaEquator), float(bnormal), *[a.ctypes.data for a in
                           (chiVal, mu, xRAM, yRAM, zRAM, bRAM, density, out, I, H, D, bz, M)])
    return I, H, D, bz, M


def hi_tail(I_cart, H_cart, HDens_cart, bZEq_cart, ScaleAt, outsideMGNP, Lz, PA, PAbn, integral_smooth, DthI, ram):
    """src/ModRamScb.f90:413-637 (scbo_hi_tail in scb_oracle.cpp); same interface as ramscb_b200.host.hI_tail, plus the
    interpolated h / I (h_Cart_interp, I_Cart_interp)."""
    lib = scb_lib()
    lib.scbo_hi_tail.argtypes = [C.c_int] * 3 + [C.c_void_p] * 9 + [C.c_int, C.c_double] + [C.c_void_p] * 12
    nR, nT, nPa = I_cart.shape
    f = lambda a: np.array(a, dtype=np.float64, order="F")
    out = {"I_cart": f(I_cart), "H_cart": f(H_cart), "HDens_cart": f(HDens_cart), "bZEq_cart": f(bZEq_cart)}
    for n in ("FNHS", "FNIS", "BOUNHS", "BOUNIS", "HDNS", "BNES"):
        out[n] = f(ram[n])
    for n in ("dIdt", "dHdt", "dIbndt"):
        out[n] = _f((nR + 1, nT, nPa))
    out["dBdt"] = _f((nR + 1, nT))
    out["h_interp"], out["I_interp"] = _f((nR, nT, nPa)), _f((nR, nT, nPa))
    sa = np.ascontiguousarray(ScaleAt, dtype=np.int32)
    om = np.asfortranarray(outsideMGNP, dtype=np.int32)
    Lz, PA, PAbn = (np.ascontiguousarray(a, dtype=np.float64) for a in (Lz, PA, PAbn))
    ptr = lambda a: a.ctypes.data
    out["gslerr"] = lib.scbo_hi_tail(nR, nT, nPa, ptr(out["I_cart"]), ptr(out["H_cart"]), ptr(out["HDens_cart"]), ptr(out["bZEq_cart"]),
                                     ptr(sa), ptr(om), ptr(Lz), ptr(PA), ptr(PAbn), 1 if integral_smooth else 0, float(DthI),
                                     *[ptr(out[n]) for n in ("FNHS", "FNIS", "BOUNHS", "BOUNIS", "HDNS", "BNES", "dIdt", "dHdt", "dIbndt",
                                                             "dBdt", "h_interp", "I_interp")])
    return out


def hi_convert_lines(x, y, z, bf, psi, alfa, Lz, MLT, nThetaEquator):
    """src/ModRamScb.f90:252-300 (hio_convert_lines); returns xRAM, yRAM, zRAM, bRAM, outsideSCB, psiRAM."""
    lib = hi_lib()
    lib.hio_convert_lines.argtypes = [C.c_int] * 6 + [C.c_void_p] * 14
    nthe, npsi, nz1 = x.shape
    f = lambda a: np.asfortranarray(a, dtype=np.float64)
    x, y, z, bf, psi, alfa = (f(a) for a in (x, y, z, bf, psi, alfa))
    Lz, MLT = (np.ascontiguousarray(a, dtype=np.float64) for a in (Lz, MLT))
    nR, nT = len(Lz) - 1, len(MLT)
    out = [_f((nthe, nR, nT)) for _ in range(4)]
    outside = np.zeros((nR, nT), dtype=np.int32, order="F")
    psiRAM = _f((nR, nT))
    lib.hio_convert_lines(nthe, npsi, nz1 - 1, nR, nT, int(nThetaEquator), *[a.ctypes.data for a in (x, y, z, bf, psi, alfa, Lz, MLT)],
                          *[a.ctypes.data for a in out], outside.ctypes.data, psiRAM.ctypes.data)
    return (*out, outside, psiRAM)


def flc_radius(x, y, z, bx, by, bz, radRaw, azimRaw, nThetaEquator, bnormal, REarth=6.4e6):
    """FLC_Radius (src/ModRamLoss.f90:176-336; hio_flc_radius): r_curvEq, zeta1Eq, zeta2Eq (nR,nT)."""
    lib = hi_lib()
    lib.hio_flc_radius.argtypes = [C.c_int] * 6 + [C.c_double] * 2 + [C.c_void_p] * 11
    lib.hio_flc_radius.restype = None
    nthe, npsi, nz1 = x.shape
    a = [np.asfortranarray(v, dtype=np.float64) for v in (x, y, z, bx, by, bz)]
    rr, az = np.ascontiguousarray(radRaw, dtype=np.float64), np.ascontiguousarray(azimRaw, dtype=np.float64)
    out = [_f((len(rr), len(az))) for _ in range(3)]
    lib.hio_flc_radius(nthe, npsi, nz1 - 1, len(rr), len(az), int(nThetaEquator), float(bnormal), float(REarth),
                       *[v.ctypes.data for v in a], rr.ctypes.data, az.ctypes.data, *[o.ctypes.data for o in out])
    return out
