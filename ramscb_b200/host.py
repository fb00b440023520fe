"""Host-side mirror of the reference's operator interface over the C ABI.

The reference calls its hot routines as ``CALL DRIFTR(S)`` etc. on module
globals (src/ModRamRun.f90:64-185).  ``RamGpu`` keeps that surface -- same
routine names, 1-based species index, same call-order requirements (DRIFTPARA
before the sweeps, CEPARA before CHAREXCHANGE/ATMOL) -- and forwards every call
through ``libramscb_gpu.so`` (include/ramscb_gpu.h) with plain host pointers, the
way the Fortran shim in ``ramscb_b200/fortran/`` does with ``c_loc``.

There is no CPU fallback: if the CUDA library is missing or no device is
present, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libramscb_gpu.so")

MODE_EXACT, MODE_FAST = 0, 1
F_WPI, F_COULOMB, F_EMIC = 1, 2, 4

_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)

PEER_BLOB_BYTES = 192
SHARD_SPECIES, SHARD_SLABS = 0, 1


class ShardPlan(C.Structure):
    """rsg_shard_plan_t (include/ramscb_gpu.h): what one rank owns in the multi-GPU RAM step"""
    _fields_ = [(n, C.c_int) for n in ("world", "rank", "policy", "s0", "ns", "G", "gidx", "g0", "l0", "nl", "b0", "nb", "per")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class RsgError(RuntimeError):
    pass


def lib():
    """Load libramscb_gpu.so (built by ``python -m ramscb_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RsgError(f"{LIB_PATH} not found: run `python -m ramscb_b200.build` (nvcc, sm_100a). "
                       "There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i, d, ll = C.c_void_p, C.c_int, C.c_double, C.c_longlong
    L.rsg_last_error.restype = C.c_char_p
    L.rsg_device_count.restype = i
    L.rsg_device_info.argtypes = [C.c_char_p, i, _ip, C.POINTER(ll), C.POINTER(ll)]
    L.rsg_ram_create.argtypes = [C.POINTER(vp), i, i, i, i, i, i]
    L.rsg_ram_destroy.argtypes = [vp]
    L.rsg_ram_set_mode.argtypes = [vp, i]
    L.rsg_ram_set_stream.argtypes = [vp, vp]
    L.rsg_ram_sync.argtypes = [vp]
    L.rsg_ram_use_graph.argtypes = [vp, i]
    L.rsg_ram_use_fused.argtypes = [vp, i]
    L.rsg_ram_set_grids.argtypes = [vp] + [vp] * 18 + [vp, vp, vp] + [d] * 6
    L.rsg_ram_set_fields.argtypes = [vp] + [vp] * 10
    L.rsg_ram_set_efield.argtypes = [vp, vp, vp, vp]
    L.rsg_ram_set_boundary.argtypes = [vp, vp]
    L.rsg_ram_set_wavelo.argtypes = [vp, vp, vp, vp, d, d]
    L.rsg_ram_set_plasmasphere.argtypes = [vp, vp]
    L.rsg_ram_set_flc_coef.argtypes = [vp, i, vp]
    L.rsg_flcscatter.argtypes = [vp, i, d, d, d, C.POINTER(C.c_longlong)]
    L.rsg_para_flc.argtypes = [vp, i, vp, vp, vp]
    L.rsg_ram_get_flc_coef.argtypes = [vp, i, vp]
    L.rsg_ram_set_diffcoef.argtypes = [vp, i, vp]
    L.rsg_geosb.argtypes = [vp, i, vp, d]
    L.rsg_ram_get_boundary.argtypes = [vp, i, vp]
    L.rsg_get_electric_field.argtypes = [vp, i, vp, vp, d, d, d, d, vp, d, vp]
    L.rsg_ram_set_wave_tables.argtypes = [vp, i, i, vp, vp, vp, vp, i, i, i, vp, vp, vp, vp, vp, vp, vp]
    L.rsg_anisch_diffcoef.argtypes = [vp, i, i, vp, i, _ip]
    L.rsg_ram_get_diffcoef.argtypes = [vp, i, vp]
    L.rsg_ram_f2_h2d.argtypes = [vp, vp, i]
    L.rsg_ram_f2_d2h.argtypes = [vp, vp, i]
    L.rsg_ram_f2_device.argtypes = [vp, i, C.POINTER(vp), C.POINTER(ll), _ip]
    for name in ("driftr", "driftp", "drifte", "driftmu", "charexchange", "atmol", "coulen"):
        getattr(L, "rsg_" + name).argtypes = [vp, i]
    for name in ("driftpara", "cepara", "wavelo", "coulpara", "coulmu"):
        getattr(L, "rsg_" + name).argtypes = [vp, i, d]
    L.rsg_driftend.argtypes = [vp]
    L.rsg_get_dtdrift.argtypes = [vp, i, _dp]
    L.rsg_wpadif.argtypes = [vp, i, d, C.POINTER(ll)]
    L.rsg_sumrc.argtypes = [vp, i, _dp, _dp]
    L.rsg_anisch.argtypes = [vp, i, vp, vp]
    L.rsg_ram_run.argtypes = [vp, d, d, d, i, _dp, vp, vp, vp, vp, vp]
    L.rsg_ram_run_host.argtypes = [vp, vp, d, d, d, i, _dp, vp, vp, vp, vp, vp]
    L.rsg_ram_part_fwd.argtypes = [vp, d, i, i, i, i, i]
    L.rsg_ram_part_all.argtypes = [vp, d, i, i, i]
    L.rsg_ram_fused_available.argtypes = [vp, i]
    L.rsg_ram_col_blocks.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.rsg_ram_fpart_planes_fwd.argtypes = [vp, d, i, i, i, i, i]
    L.rsg_ram_fpart_columns.argtypes = [vp, d, i, i, i, i, i]
    L.rsg_ram_fpart_planes_rev.argtypes = [vp, i, i, i, i]
    L.rsg_ram_results_device.argtypes = [vp, C.POINTER(C.c_void_p), C.POINTER(C.c_longlong), C.POINTER(C.c_void_p), C.POINTER(C.c_longlong)]
    L.rsg_ram_part_mid.argtypes = [vp, d, i, i, i, i, i]
    L.rsg_ram_part_rev.argtypes = [vp, i, i, i, i]
    L.rsg_ram_part_results.argtypes = [vp, i, i, vp, vp, vp, vp]
    L.rsg_ram_flux_d2h.argtypes = [vp, vp]
    L.rsg_ram_launch_count.argtypes = [vp]
    L.rsg_ram_launch_count.restype = ll
    L.rsg_ram_timer_begin.argtypes = [vp]
    L.rsg_ram_timer_end.argtypes = [vp, _dp]
    L.rsg_ram_profile.argtypes = [vp, i]
    L.rsg_ram_profile_get.argtypes = [vp, i, C.c_char_p, i, _dp, C.POINTER(ll)]
    L.rsg_host_register.argtypes = [vp, ll]
    L.rsg_host_unregister.argtypes = [vp]
    L.rsg_shard_plan.argtypes = [i, i, i, i, i, i, C.POINTER(ShardPlan)]
    L.rsg_ram_shard_info.argtypes = [vp, C.POINTER(ShardPlan)]
    L.rsg_ram_peer_export.argtypes = [vp, vp]
    L.rsg_ram_peer_attach.argtypes = [vp, i, i, i, vp]
    L.rsg_ram_peer_attach_local.argtypes = [vp, i, i, i, C.POINTER(vp)]
    L.rsg_ram_peer_detach.argtypes = [vp]
    L.rsg_ram_run_sharded.argtypes = [vp, d, d, d, i, _dp, vp, vp, vp, vp, vp]
    L.rsg_ram_run_sharded_enqueue.argtypes = [vp, d, d, i]
    L.rsg_ram_run_sharded_collect.argtypes = [vp, d, _dp, vp, vp, vp, vp, vp]
    L.rsg_ram_f2_h2d_shard.argtypes = [vp, vp]
    L.rsg_ram_f2_d2h_shard.argtypes = [vp, vp]
    _lib = L
    return L


def _ck(rc):
    if rc != 0:
        raise RsgError(f"rsg status {rc}: {lib().rsg_last_error().decode()}")


def _p(a, dtype=np.float64):
    """Pointer to a Fortran-contiguous array of the right dtype (no copy made)."""
    if a.dtype != dtype or not (a.flags.f_contiguous or a.ndim <= 1 and a.flags.c_contiguous):
        raise ValueError("array must be Fortran-contiguous " + str(dtype))
    return a.ctypes.data


def host_register(a):
    """Pin a numpy array in place (cudaHostRegister)."""
    _ck(lib().rsg_host_register(a.ctypes.data, a.nbytes))


def host_unregister(a):
    _ck(lib().rsg_host_unregister(a.ctypes.data))


def device_count() -> int:
    return lib().rsg_device_count()


def shard_plan(world, rank, policy, nS, NPA, P):
    """The decomposition of the sharded RAM step as the library computes it (no device needed)."""
    p = ShardPlan()
    _ck(lib().rsg_shard_plan(world, rank, policy, nS, NPA, P, C.byref(p)))
    return p


def device_info():
    name = C.create_string_buffer(256)
    sm = C.c_int()
    l2 = C.c_longlong()
    hbm = C.c_longlong()
    _ck(lib().rsg_device_info(name, 256, C.byref(sm), C.byref(l2), C.byref(hbm)))
    return {"name": name.value.decode(), "sm_count": sm.value, "l2_bytes": l2.value, "hbm_bytes": hbm.value}


class RamGpu:
    """Device-resident RAM state + the reference's operator names."""

    def __init__(self, g, device: int = -1, mode: int = MODE_EXACT):
        self.L = lib()
        self.g = g
        self.h = C.c_void_p()
        _ck(self.L.rsg_ram_create(C.byref(self.h), g.nS, g.NR, g.NT, g.NE, g.NPA, device))
        self._keep = []
        self.set_grids(g)
        if mode != MODE_EXACT:
            self.set_mode(mode)

    # ---- configuration --------------------------------------------------------
    def set_mode(self, mode):
        _ck(self.L.rsg_ram_set_mode(self.h, mode))

    def set_stream(self, stream_ptr):
        _ck(self.L.rsg_ram_set_stream(self.h, C.c_void_p(stream_ptr) if stream_ptr else None))

    def sync(self):
        _ck(self.L.rsg_ram_sync(self.h))

    def use_fused(self, on=True, wpadif=True):
        """Fused shared-memory kernels of the FAST step; ``wpadif=False`` keeps WPADIF (WPI / EMIC
        flags) as its own kernel, which then takes the whole step to the one-kernel-per-operator path."""
        _ck(self.L.rsg_ram_use_fused(self.h, (1 if wpadif else 3) if on else 0))

    def use_graph(self, on=True):
        _ck(self.L.rsg_ram_use_graph(self.h, 1 if on else 0))

    def set_grids(self, g, BetaLim=1.5, FracCFL=0.8):
        f = lambda a: _p(np.asfortranarray(a, dtype=np.float64))
        arrs = [np.asfortranarray(getattr(g, n), dtype=np.float64) for n in
                ("RLZ", "LZ", "EKEV", "WE", "DE", "EBND", "MU", "WMU", "DMU", "UPA", "GREL", "GRBND", "V", "VBND",
                 "EPP", "ERNH", "RMAS", "FFACTOR")]
        ints = [np.ascontiguousarray(getattr(g, n), dtype=np.int32) for n in ("QS", "kind", "khi")]
        _ck(self.L.rsg_ram_set_grids(self.h, *[a.ctypes.data for a in arrs], *[a.ctypes.data for a in ints],
                                     g.MDR, g.DPHI, g.CONF1, g.CONF2, BetaLim, FracCFL))

    def set_fields(self, inp):
        out = np.asfortranarray(inp.outsideMGNP, dtype=np.int32)
        _ck(self.L.rsg_ram_set_fields(self.h, _p(inp.BNES), _p(inp.dBdt), _p(inp.FNHS), _p(inp.FNIS), _p(inp.BOUNHS),
                                      _p(inp.BOUNIS), _p(inp.HDNS), _p(inp.dIdt), _p(inp.dIbndt), out.ctypes.data))

    def set_efield(self, VT, EIR, EIP):
        _ck(self.L.rsg_ram_set_efield(self.h, _p(VT), _p(EIR), _p(EIP)))

    def set_boundary(self, FGEOS):
        _ck(self.L.rsg_ram_set_boundary(self.h, _p(FGEOS)))

    def set_wavelo(self, W1, W2, W3, Kp, Kpmax12):
        _ck(self.L.rsg_ram_set_wavelo(self.h, _p(W1), _p(W2), _p(W3), Kp, Kpmax12))

    def set_plasmasphere(self, NECR):
        _ck(self.L.rsg_ram_set_plasmasphere(self.h, _p(NECR)))

    def set_diffcoef(self, which, D):
        _ck(self.L.rsg_ram_set_diffcoef(self.h, which, _p(D)))

    def GEOSB(self, S, FluxLanl, s_comp=1.0):
        """GEOSB (src/ModRamBoundary.f90:241-319, boundary LANL) on the device from the geosynchronous flux (NT,NE)"""
        _ck(self.L.rsg_geosb(self.h, S, _p(np.asfortranarray(FluxLanl, dtype=np.float64)), s_comp))

    def get_boundary(self, S):
        g = self.g
        out = np.zeros((g.NT, g.NE, g.NPA), order="F")
        _ck(self.L.rsg_ram_get_boundary(self.h, S, _p(out)))
        return out

    def get_electric_field(self, vols, VTOL=None, VTN=None, t=0.0, TOLV=0.0, DtEfi=1.0, Kp=0.0, PHI=None, PHIOFS=0.0):
        """get_electric_field (src/ModRamEField.f90:14-63) on the device; returns VT(NR+1,NT)"""
        g = self.g
        VT = np.zeros((g.NR + 1, g.NT), order="F")
        a = _p(np.asfortranarray(VTOL, dtype=np.float64)) if VTOL is not None else None
        b = _p(np.asfortranarray(VTN, dtype=np.float64)) if VTN is not None else None
        ph = np.ascontiguousarray(PHI, dtype=np.float64) if PHI is not None else None
        _ck(self.L.rsg_get_electric_field(self.h, 1 if vols else 0, a, b, t, TOLV, DtEfi, Kp, ph.ctypes.data if ph is not None else None,
                                          PHIOFS, _p(VT)))
        return VT

    def set_wave_tables(self, t, use_bas=True):
        """the tabulated diffusion coefficients (dict of synthetic.synthetic_wave_tables / the reference's start-up files)"""
        f = lambda n: np.asfortranarray(t[n], dtype=np.float64)
        a = {n: f(n) for n in ("ENOR", "fpofc", "NDAAJ", "CDAAR" if use_bas else "BDAAR", "EKEV_emic", "fp2c_emic", "Daa_emic_h",
                               "Daa_emic_he", "Ihs_emic", "Ihes_emic", "PAbn")}
        self._keep.append(a)
        _ck(self.L.rsg_ram_set_wave_tables(self.h, int(t["ENG"]), int(t["NCF"]), a["ENOR"].ctypes.data, a["fpofc"].ctypes.data,
                                           a["NDAAJ"].ctypes.data, a["CDAAR" if use_bas else "BDAAR"].ctypes.data, 1 if use_bas else 0,
                                           int(t["ENG_emic"]), int(t["NCF_emic"]), a["EKEV_emic"].ctypes.data, a["fp2c_emic"].ctypes.data,
                                           a["Daa_emic_h"].ctypes.data, a["Daa_emic_he"].ctypes.data, a["Ihs_emic"].ctypes.data,
                                           a["Ihes_emic"].ctypes.data, a["PAbn"].ctypes.data))

    def ANISCH_diffcoef(self, S, flags, XNE, AE=0):
        """second half of ANISCH (src/ModRamRun.f90:422-605): rebuild ATAW/ATAC (electrons, WPI) or ATAW_emic_h/_he (H+, EMIC)
        on the device; returns the reference's GSLerr count"""
        err = C.c_int()
        _ck(self.L.rsg_anisch_diffcoef(self.h, S, flags, _p(np.asfortranarray(XNE, dtype=np.float64)), int(AE), C.byref(err)))
        return err.value

    def get_diffcoef(self, which):
        g = self.g
        out = np.zeros((g.NR, g.NT, g.NE, g.NPA), order="F")
        _ck(self.L.rsg_ram_get_diffcoef(self.h, which, _p(out)))
        return out

    def set_inputs(self, inp):
        """Everything a RamInputs carries, F2 included."""
        self.set_fields(inp)
        self.set_efield(inp.VT, inp.EIR, inp.EIP)
        self.set_boundary(inp.FGEOS)
        self.set_wavelo(inp.WALOS1, inp.WALOS2, inp.WALOS3, inp.Kp, inp.Kpmax12)
        self.set_plasmasphere(inp.NECR)
        self.f2_h2d(inp.F2)

    # ---- F2 -------------------------------------------------------------------
    def f2_h2d(self, F2, S=0):
        _ck(self.L.rsg_ram_f2_h2d(self.h, _p(F2), S))

    def f2_d2h(self, F2=None, S=0):
        g = self.g
        if F2 is None:
            F2 = np.zeros((g.nS, g.NR, g.NT, g.NE, g.NPA), order="F")
        _ck(self.L.rsg_ram_f2_d2h(self.h, _p(F2), S))
        return F2

    def f2_device(self, S):
        ptr, n, pp = C.c_void_p(), C.c_longlong(), C.c_int()
        _ck(self.L.rsg_ram_f2_device(self.h, S, C.byref(ptr), C.byref(n), C.byref(pp)))
        return ptr.value, n.value, pp.value

    # ---- the reference's operator names (1-based species) ----------------------
    def DRIFTPARA(self, S, DTs): _ck(self.L.rsg_driftpara(self.h, S, DTs))
    def DRIFTR(self, S): _ck(self.L.rsg_driftr(self.h, S))
    def DRIFTP(self, S): _ck(self.L.rsg_driftp(self.h, S))
    def DRIFTE(self, S): _ck(self.L.rsg_drifte(self.h, S))
    def DRIFTMU(self, S): _ck(self.L.rsg_driftmu(self.h, S))
    def DRIFTEND(self): _ck(self.L.rsg_driftend(self.h))
    def CEPARA(self, S, DTs): _ck(self.L.rsg_cepara(self.h, S, DTs))
    def CHAREXCHANGE(self, S): _ck(self.L.rsg_charexchange(self.h, S))
    def ATMOL(self, S): _ck(self.L.rsg_atmol(self.h, S))
    def WAVELO(self, S, DTs): _ck(self.L.rsg_wavelo(self.h, S, DTs))
    def set_flc_coef(self, S, D):
        _ck(self.L.rsg_ram_set_flc_coef(self.h, S, _p(np.asfortranarray(D, dtype=np.float64))))

    def PARA_FLC(self, S, r_curvEq, zeta1Eq, zeta2Eq):
        a = [np.asfortranarray(v, dtype=np.float64) for v in (r_curvEq, zeta1Eq, zeta2Eq)]
        _ck(self.L.rsg_para_flc(self.h, S, _p(a[0]), _p(a[1]), _p(a[2])))

    def get_flc_coef(self, S):
        g = self.g
        out = np.zeros((g.NR, g.NT, g.NE, g.NPA), order="F")
        _ck(self.L.rsg_ram_get_flc_coef(self.h, S, _p(out)))
        return out

    def FLCscatter(self, S, DTs, T, Dt_bc=300.0):
        nv = C.c_longlong()
        _ck(self.L.rsg_flcscatter(self.h, S, DTs, T, Dt_bc, C.byref(nv)))
        return nv.value

    def COULPARA(self, S, DTs): _ck(self.L.rsg_coulpara(self.h, S, DTs))
    def COULEN(self, S): _ck(self.L.rsg_coulen(self.h, S))
    def COULMU(self, S, T=0.0): _ck(self.L.rsg_coulmu(self.h, S, T))

    def WPADIF(self, S, DTs):
        nv = C.c_longlong()
        _ck(self.L.rsg_wpadif(self.h, S, DTs, C.byref(nv)))
        return nv.value

    def dtdrift(self, S):
        out = (C.c_double * 4)()
        _ck(self.L.rsg_get_dtdrift(self.h, S, out))
        return np.array(out[:])

    def SUMRC(self, S):
        a, b = C.c_double(), C.c_double()
        _ck(self.L.rsg_sumrc(self.h, S, C.byref(a), C.byref(b)))
        return a.value, b.value

    def ANISCH(self, S):
        g = self.g
        pper = np.zeros((g.NR, g.NT), order="F")
        ppar = np.zeros((g.NR, g.NT), order="F")
        _ck(self.L.rsg_anisch(self.h, S, _p(pper), _p(ppar)))
        return pper, ppar

    def ram_run(self, DTs, DtsMin=1.0, T=0.0, flags=0):
        g = self.g
        dtn = C.c_double()
        out = {
            "DtDrift": np.zeros((4, g.nS), order="F"),
            "losses": np.zeros((6, g.nS), order="F"),
            "SETRC": np.zeros(g.nS),
            "PPERT": np.zeros((g.nS, g.NR, g.NT), order="F"),
            "PPART": np.zeros((g.nS, g.NR, g.NT), order="F"),
        }
        _ck(self.L.rsg_ram_run(self.h, DTs, DtsMin, T, flags, C.byref(dtn), _p(out["DtDrift"]), _p(out["losses"]),
                               _p(out["SETRC"]), _p(out["PPERT"]), _p(out["PPART"])))
        out["DtsNext"] = dtn.value
        return out

    def ram_run_host(self, F2, DTs, DtsMin=1.0, T=0.0, flags=0):
        """ram_run with F2 (the host array, Fortran order (nS,NR,NT,NE,NPA), updated in place) going up and coming back
        inside the call, pipelined over chunks of pitch angles (rsg_ram_run_host)."""
        g = self.g
        assert F2.flags.f_contiguous and F2.dtype == np.float64 and F2.shape == (g.nS, g.NR, g.NT, g.NE, g.NPA)
        dtn = C.c_double()
        out = {
            "DtDrift": np.zeros((4, g.nS), order="F"),
            "losses": np.zeros((6, g.nS), order="F"),
            "SETRC": np.zeros(g.nS),
            "PPERT": np.zeros((g.nS, g.NR, g.NT), order="F"),
            "PPART": np.zeros((g.nS, g.NR, g.NT), order="F"),
        }
        _ck(self.L.rsg_ram_run_host(self.h, F2.ctypes.data, DTs, DtsMin, T, flags, C.byref(dtn), _p(out["DtDrift"]), _p(out["losses"]),
                                    _p(out["SETRC"]), _p(out["PPERT"]), _p(out["PPART"])))
        out["DtsNext"] = dtn.value
        return out

    # ---- multi-GPU parts (include/ramscb_gpu.h: rsg_ram_part_*) ---------------------
    def part_fwd(self, DTs, flags, s0, ns, l0, nl):
        _ck(self.L.rsg_ram_part_fwd(self.h, DTs, flags, s0, ns, l0, nl))

    def results_device(self):
        """(res_ptr, words per species, pp_ptr, doubles per species) of the device result blocks."""
        r, rn, q, qn = C.c_void_p(), C.c_longlong(), C.c_void_p(), C.c_longlong()
        _ck(self.L.rsg_ram_results_device(self.h, C.byref(r), C.byref(rn), C.byref(q), C.byref(qn)))
        return r.value, rn.value, q.value, qn.value

    def fused_available(self, flags=0):
        return bool(self.L.rsg_ram_fused_available(self.h, flags))

    def col_blocks(self):
        n, per = C.c_int(), C.c_int()
        _ck(self.L.rsg_ram_col_blocks(self.h, C.byref(n), C.byref(per)))
        return n.value, per.value

    def fpart_planes_fwd(self, DTs, flags, s0, ns, l0, nl):
        _ck(self.L.rsg_ram_fpart_planes_fwd(self.h, DTs, flags, s0, ns, l0, nl))

    def fpart_columns(self, DTs, flags, s0, ns, b0, nb):
        _ck(self.L.rsg_ram_fpart_columns(self.h, DTs, flags, s0, ns, b0, nb))

    def fpart_planes_rev(self, s0, ns, l0, nl):
        _ck(self.L.rsg_ram_fpart_planes_rev(self.h, s0, ns, l0, nl))

    def part_all(self, DTs, flags, s0, ns):
        _ck(self.L.rsg_ram_part_all(self.h, DTs, flags, s0, ns))

    def part_mid(self, DTs, flags, s0, ns, k0, nk):
        _ck(self.L.rsg_ram_part_mid(self.h, DTs, flags, s0, ns, k0, nk))

    def part_rev(self, s0, ns, l0, nl):
        _ck(self.L.rsg_ram_part_rev(self.h, s0, ns, l0, nl))

    def part_results(self, s0, ns):
        g = self.g
        dt = np.zeros((4, ns), order="F")
        mom = np.zeros((14, ns), order="F")
        pper = np.zeros((g.NR, g.NT, ns), order="F")
        ppar = np.zeros((g.NR, g.NT, ns), order="F")
        _ck(self.L.rsg_ram_part_results(self.h, s0, ns, _p(dt), _p(mom), _p(pper), _p(ppar)))
        return dt, mom, pper, ppar

    # ---- multi-GPU inside the library (include/ramscb_gpu.h: rsg_ram_peer_*, rsg_ram_run_sharded) --------
    def peer_export(self):
        blob = np.zeros(PEER_BLOB_BYTES, dtype=np.uint8)
        _ck(self.L.rsg_ram_peer_export(self.h, blob.ctypes.data))
        return blob

    def peer_attach(self, rank, world, policy, blobs):
        blobs = np.ascontiguousarray(blobs, dtype=np.uint8).reshape(world, PEER_BLOB_BYTES)
        _ck(self.L.rsg_ram_peer_attach(self.h, rank, world, policy, blobs.ctypes.data))

    def peer_attach_local(self, rank, peers, policy=SHARD_SPECIES):
        arr = (C.c_void_p * len(peers))(*[p.h for p in peers])
        _ck(self.L.rsg_ram_peer_attach_local(self.h, rank, len(peers), policy, arr))

    def peer_detach(self):
        _ck(self.L.rsg_ram_peer_detach(self.h))

    def shard_info(self):
        p = ShardPlan()
        _ck(self.L.rsg_ram_shard_info(self.h, C.byref(p)))
        return p

    def _step_outputs(self):
        g = self.g
        return {"DtDrift": np.zeros((4, g.nS), order="F"), "losses": np.zeros((6, g.nS), order="F"), "SETRC": np.zeros(g.nS),
                "PPERT": np.zeros((g.nS, g.NR, g.NT), order="F"), "PPART": np.zeros((g.nS, g.NR, g.NT), order="F")}

    def run_sharded(self, DTs, DtsMin=1.0, T=0.0, flags=0):
        """ram_run on this rank's share; every rank returns the results of ALL species (like ``ram_run``)."""
        dtn, out = C.c_double(), self._step_outputs()
        _ck(self.L.rsg_ram_run_sharded(self.h, DTs, DtsMin, T, flags, C.byref(dtn), _p(out["DtDrift"]), _p(out["losses"]),
                                       _p(out["SETRC"]), _p(out["PPERT"]), _p(out["PPART"])))
        out["DtsNext"] = dtn.value
        return out

    def run_sharded_enqueue(self, DTs, T=0.0, flags=0):
        _ck(self.L.rsg_ram_run_sharded_enqueue(self.h, DTs, T, flags))

    def run_sharded_collect(self, DtsMin=1.0):
        dtn, out = C.c_double(), self._step_outputs()
        _ck(self.L.rsg_ram_run_sharded_collect(self.h, DtsMin, C.byref(dtn), _p(out["DtDrift"]), _p(out["losses"]), _p(out["SETRC"]),
                                               _p(out["PPERT"]), _p(out["PPART"])))
        out["DtsNext"] = dtn.value
        return out

    def f2_h2d_shard(self, F2):
        _ck(self.L.rsg_ram_f2_h2d_shard(self.h, _p(F2)))

    def f2_d2h_shard(self, F2):
        _ck(self.L.rsg_ram_f2_d2h_shard(self.h, _p(F2)))
        return F2

    def flux_d2h(self):
        g = self.g
        FLUX = np.zeros((g.nS, g.NR, g.NT, g.NE, g.NPA), order="F")
        _ck(self.L.rsg_ram_flux_d2h(self.h, _p(FLUX)))
        return FLUX

    def launch_count(self):
        return self.L.rsg_ram_launch_count(self.h)

    def profile(self, on=True):
        _ck(self.L.rsg_ram_profile(self.h, 1 if on else 0))

    def profile_get(self):
        """{stage: (total_ms, count)} accumulated by rsg_ram_run since profile(True)."""
        out = {}
        idx = 0
        while True:
            name = C.create_string_buffer(64)
            ms, cnt = C.c_double(), C.c_longlong()
            if self.L.rsg_ram_profile_get(self.h, idx, name, 64, C.byref(ms), C.byref(cnt)) != 0:
                break
            out[name.value.decode()] = (ms.value, cnt.value)
            idx += 1
        return out

    def timer_begin(self):
        _ck(self.L.rsg_ram_timer_begin(self.h))

    def timer_end(self):
        ms = C.c_double()
        _ck(self.L.rsg_ram_timer_end(self.h, C.byref(ms)))
        return ms.value

    def close(self):
        if self.h:
            self.L.rsg_ram_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# =============================================================================
# SCB: mirror of ModScbCompute / ModScbEquation / ModScbEuler over the C ABI
# =============================================================================
SOR_LEX, SOR_COLOR4 = 0, 1
_scb_ready = False


class ScbRunParams(C.Structure):
    """rsg_scb_run_params (include/ramscb_gpu.h); defaults = the reference's (ModScbParams.f90, ModScbMain.f90)"""
    _fields_ = [(n, C.c_double) for n in ("InConAlpha", "InConPsi", "blendInitial", "blendMin", "blendMax", "damp",
                                          "decreaseConvAlpha", "decreaseConvPsi")] + \
               [(n, C.c_int) for n in ("nimax", "theChange", "psiChange", "numit", "MinSCBIterations", "ordering", "iLossCone",
                                       "iReduceAnisotropy")]


class ScbRunResult(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("iterations", "iConvGlobal", "SORFail", "nisaveAlpha", "nisavePsi", "blendRetries")] + \
               [(n, C.c_double) for n in ("blendAlpha", "blendPsi", "errorAlpha", "errorPsi", "sumbAlpha", "sumdbAlpha", "sumbPsi",
                                          "sumdbPsi", "normDiffStart", "normJxBStart", "normGradPStart", "normDiff", "normJxB",
                                          "normGradP")]


SCB_PRESSURE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                              C.POINTER(C.c_double), C.POINTER(C.c_double))


def _scb_lib():
    global _scb_ready
    L = lib()
    if not _scb_ready:
        vp, i, d, ll = C.c_void_p, C.c_int, C.c_double, C.c_longlong
        L.rsg_scb_last_error.restype = C.c_char_p
        L.rsg_scb_create.argtypes = [C.POINTER(vp), i, i, i, i]
        L.rsg_scb_destroy.argtypes = [vp]
        L.rsg_scb_set_grid.argtypes = [vp] * 6
        L.rsg_scb_set_geometry.argtypes = [vp] * 4
        L.rsg_scb_set_pressure.argtypes = [vp, i] + [vp] * 15
        L.rsg_scb_flc_radius.argtypes = [vp, i, i, vp, vp, d, vp, vp, vp]
        L.rsg_hi_create.argtypes = [C.POINTER(vp), i, i, i, i, i, i, i, i, d] + [vp] * 6
        L.rsg_hi_destroy.argtypes = [vp]
        L.rsg_hi_destroy.restype = None
        L.rsg_hi_set_ram_fields.argtypes = [vp] * 8
        L.rsg_hi_convert.argtypes = [vp] * 8 + [vp, _ip]
        L.rsg_hi_set_line.argtypes = [vp, i, i, vp, vp, vp, vp]
        L.rsg_hi_finish.argtypes = [vp, vp, vp, vp, i, d, _ip]
        L.rsg_computehI.argtypes = [vp] * 8 + [i, d, _ip]
        L.rsg_hi_get.argtypes = [vp, C.c_char_p, vp]
        L.rsg_hi_get_int.argtypes = [vp, i, vp]
        L.rsg_hi_last_ms.argtypes = [vp]
        L.rsg_hi_last_ms.restype = d
        L.rsg_hi_launch_count.argtypes = [vp]
        L.rsg_hi_launch_count.restype = ll
        L.rsg_hi_device_fields.argtypes = [vp, vp, vp]
        L.rsg_ram_set_fields_device.argtypes = [vp, vp, vp]
        L.rsg_scb_set_ram_pressure.argtypes = [vp, i, i, i, vp, vp, vp, vp, vp, i, i, i]
        L.rsg_scb_get_ram_pressure.argtypes = [vp, _ip, _ip, vp, vp, vp, vp]
        L.rsg_scb_pressure_front.argtypes = [vp, i, i, vp, vp]
        L.rsg_scb_set_field.argtypes = [vp, C.c_char_p, vp]
        L.rsg_scb_get_field.argtypes = [vp, C.c_char_p, vp]
        L.rsg_scb_field_size.argtypes = [vp, C.c_char_p, C.POINTER(ll)]
        L.rsg_scb_bandjacob.argtypes = [vp, _ip]
        for n in ("metrica", "metric", "newk", "newj"):
            getattr(L, "rsg_scb_" + n).argtypes = [vp]
        for n in ("iterate_alpha", "iterate_psi"):
            getattr(L, "rsg_scb_" + n).argtypes = [vp, d, i, i, i, i, _ip, _dp, _dp, _dp, _ip, vp]
        L.rsg_scb_convergence.argtypes = [vp, _dp, _dp, _dp, _ip]
        L.rsg_scb_derivs.argtypes = [vp] * 5
        L.rsg_scb_set_map_targets.argtypes = [vp, vp, vp, vp]
        for n in ("map_alpha", "map_psi", "map_theta"):
            getattr(L, "rsg_scb_" + n).argtypes = [vp, _ip]
        L.rsg_scb_pressure_aniso.argtypes = [vp, vp, vp, i, i]
        L.rsg_scb_snapshot.argtypes = [vp, C.c_char_p, i]
        L.rsg_scb_restore.argtypes = [vp, C.c_char_p, i]
        L.rsg_scb_blend.argtypes = [vp, C.c_char_p, i, i, d]
        L.rsg_scb_min_jacobian.argtypes = [vp, _dp]
        L.rsg_scb_last_ms.argtypes = [vp]
        L.rsg_scb_last_ms.restype = d
        L.rsg_scb_use_cluster.argtypes = [vp, i]
        L.rsg_scb_iterate_part.argtypes = [vp, i, d, i, i, i, i, i, i]
        L.rsg_scb_iterate_finish.argtypes = [vp, i, i, i, C.POINTER(C.c_int), C.POINTER(d), C.POINTER(d), C.POINTER(d),
                                             C.POINTER(C.c_int), vp]
        L.rsg_scb_field_device.argtypes = [vp, C.c_char_p, C.POINTER(vp), C.POINTER(C.c_longlong)]
        L.rsg_scb_set_stream.argtypes = [vp, vp]
        L.rsg_scb_run.argtypes = [vp, C.POINTER(ScbRunParams), SCB_PRESSURE_FN, vp, C.POINTER(ScbRunResult)]
        L.rsg_scb_zsolve_begin.argtypes = [vp, d, i, i, i, i, i]
        L.rsg_scb_zsolve_half.argtypes = [vp, i]
        L.rsg_scb_zsolve_state_device.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_longlong)]
        L.rsg_scb_zsolve_commit.argtypes = [vp]
        L.rsg_scb_zsolve_pending.argtypes = [vp, C.POINTER(C.c_int)]
        L.rsg_scb_last_cluster.argtypes = [vp]
        L.rsg_scb_launch_count.argtypes = [vp]
        L.rsg_scb_launch_count.restype = ll
        L.rsg_hI_integrals.argtypes = [i, i, i, i, i, i, d] + [vp] * 12 + [_dp]
        L.rsg_hI_convert_lines.argtypes = [i] * 7 + [vp] * 13 + [_dp]
        L.rsg_hI_tail.argtypes = [i, i, i, i] + [vp] * 9 + [i, d] + [vp] * 10 + [_ip, _dp]
        _scb_ready = True
    return L


def _sck(rc):
    if rc != 0:
        raise RsgError(f"rsg_scb status {rc}: {_scb_lib().rsg_scb_last_error().decode()}")


class ScbGpu:
    """Device-resident SCB state with the reference's routine names
    (computeBandJacob, metrica, metric, newk, newj, iterateAlpha, iteratePsi,
    Compute_convergence -- all argument-less in the reference)."""

    def __init__(self, inp, device: int = -1):
        self.L = _scb_lib()
        self.inp = inp
        self.nthe, self.npsi, self.nzeta = inp.nthe, inp.npsi, inp.nzeta
        self.h = C.c_void_p()
        _sck(self.L.rsg_scb_create(C.byref(self.h), inp.nthe, inp.npsi, inp.nzeta, device))
        c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        g = [c(inp.thetaVal), c(inp.rhoVal), c(inp.zetaVal), c(inp.f), c(inp.fzet)]
        _sck(self.L.rsg_scb_set_grid(self.h, *[a.ctypes.data for a in g]))
        _sck(self.L.rsg_scb_set_geometry(self.h, _p(inp.x), _p(inp.y), _p(inp.z)))
        self.set_pressure(inp)
        self.set_field("alfa", inp.alfa)
        self.set_field("psi", inp.psi)
        if getattr(inp, "chiVal", None) is not None:
            self.set_map_targets(inp.alphaVal, inp.psiVal, inp.chiVal)

    def set_pressure(self, inp):
        names = ("pper", "ppar", "sigma", "dPPerdTheta", "dPPerdRho", "dPPerdZeta", "dBsqdTheta", "dBsqdRho", "dBsqdZeta",
                 "dPPerdPsi", "dPPerdAlpha", "dBsqdPsi", "dBsqdAlpha", "dPdAlpha", "dPdPsi")
        _sck(self.L.rsg_scb_set_pressure(self.h, inp.isotropy, *[_p(getattr(inp, n)) for n in names]))

    def set_geometry(self, x, y, z):
        _sck(self.L.rsg_scb_set_geometry(self.h, _p(x), _p(y), _p(z)))

    def set_field(self, name, a):
        _sck(self.L.rsg_scb_set_field(self.h, name.encode(), _p(np.asfortranarray(a, dtype=np.float64))))

    def get_field(self, name):
        n = C.c_longlong()
        _sck(self.L.rsg_scb_field_size(self.h, name.encode(), C.byref(n)))
        plane = self.nthe * self.npsi
        if n.value % plane == 0 and n.value // plane in (self.nzeta, self.nzeta + 1):
            out = np.zeros((self.nthe, self.npsi, n.value // plane), order="F")
        else:
            out = np.zeros(n.value)
        _sck(self.L.rsg_scb_get_field(self.h, name.encode(), out.ctypes.data))
        return out

    def computeBandJacob(self):
        f = C.c_int()
        _sck(self.L.rsg_scb_bandjacob(self.h, C.byref(f)))
        return f.value

    def set_map_targets(self, alphaVal, psiVal, chiVal):
        a = [np.ascontiguousarray(v, dtype=np.float64) for v in (alphaVal, psiVal, chiVal)]
        _sck(self.L.rsg_scb_set_map_targets(self.h, *[v.ctypes.data for v in a]))

    def _map(self, fn):
        f = C.c_int()
        _sck(fn(self.h, C.byref(f)))
        return f.value

    def pressure_aniso(self, pperEq, pparEq, iLossCone=1, iReduceAnisotropy=0):
        """Tail of `pressure` (src/ModScbRun.f90:1087-1175) from the normalised equatorial pressures."""
        a, b = (np.asfortranarray(v, dtype=np.float64) for v in (pperEq, pparEq))
        _sck(self.L.rsg_scb_pressure_aniso(self.h, _p(a), _p(b), iLossCone, iReduceAnisotropy))

    # glue of the outer iteration (src/ModScbRun.f90:232-262): device snapshots, blend, Jacobian sign test
    def snapshot(self, name, slot): _sck(self.L.rsg_scb_snapshot(self.h, name.encode(), slot))
    def restore(self, name, slot): _sck(self.L.rsg_scb_restore(self.h, name.encode(), slot))
    def blend(self, name, slot_new, slot_sav, w): _sck(self.L.rsg_scb_blend(self.h, name.encode(), slot_new, slot_sav, float(w)))

    def min_jacobian(self):
        v = C.c_double()
        _sck(self.L.rsg_scb_min_jacobian(self.h, C.byref(v)))
        return v.value

    def mapAlpha(self): return self._map(self.L.rsg_scb_map_alpha)     # src/ModScbEuler.f90:97
    def mapPsi(self): return self._map(self.L.rsg_scb_map_psi)         # :403
    def mapTheta(self): return self._map(self.L.rsg_scb_map_theta)     # :15

    def metrica(self): _sck(self.L.rsg_scb_metrica(self.h))
    def metric(self): _sck(self.L.rsg_scb_metric(self.h))
    def newk(self): _sck(self.L.rsg_scb_newk(self.h))
    def newj(self): _sck(self.L.rsg_scb_newj(self.h))

    def _iterate(self, fn, tol, nimax, theChange, psiChange, ordering, n):
        nisave, fail = C.c_int(), C.c_int()
        sumb, sumdb, diffmx = C.c_double(), C.c_double(), C.c_double()
        ni = np.zeros(n, dtype=np.int32)
        _sck(fn(self.h, tol, nimax, theChange, psiChange, ordering, C.byref(nisave), C.byref(sumb), C.byref(sumdb),
                C.byref(diffmx), C.byref(fail), ni.ctypes.data))
        return {"nisave": nisave.value, "sumb": sumb.value, "sumdb": sumdb.value, "diffmx": diffmx.value,
                "SORFail": fail.value, "ni": ni, "ms": self.last_ms()}

    # ---- multi-GPU: sub-problem ranges (include/ramscb_gpu.h: rsg_scb_iterate_part/finish) ----
    def iterate_part(self, alpha, tol, sub0, nsub, nimax=5001, theChange=4, psiChange=0, ordering=SOR_COLOR4):
        _sck(self.L.rsg_scb_iterate_part(self.h, 1 if alpha else 0, tol, nimax, theChange, psiChange, ordering, sub0, nsub))

    def iterate_finish(self, alpha, theChange=4, psiChange=0):
        nisave, fail = C.c_int(), C.c_int()
        sumb, sumdb, diffmx = C.c_double(), C.c_double(), C.c_double()
        ni = np.zeros(self.npsi if alpha else self.nzeta, dtype=np.int32)
        _sck(self.L.rsg_scb_iterate_finish(self.h, 1 if alpha else 0, theChange, psiChange, C.byref(nisave), C.byref(sumb),
                                           C.byref(sumdb), C.byref(diffmx), C.byref(fail), ni.ctypes.data))
        return {"nisave": nisave.value, "sumb": sumb.value, "sumdb": sumdb.value, "diffmx": diffmx.value,
                "SORFail": fail.value, "ni": ni, "ms": self.last_ms()}

    def FLC_Radius(self, radRaw, azimRaw, REarth=6.4e6):
        """FLC_Radius (src/ModRamLoss.f90:176-336) from the resident geometry / field of the last computeBandJacob:
        r_curvEq, zeta1Eq, zeta2Eq (nR,nT), the inputs of RamGpu.PARA_FLC"""
        rr, az = np.ascontiguousarray(radRaw, dtype=np.float64), np.ascontiguousarray(azimRaw, dtype=np.float64)
        out = [np.zeros((len(rr), len(az)), order="F") for _ in range(3)]
        _sck(self.L.rsg_scb_flc_radius(self.h, len(rr), len(az), rr.ctypes.data, az.ctypes.data, REarth, *[o.ctypes.data for o in out]))
        return out

    PRESS_MODES = {"SKD": 0, "ROE": 1, "EXT": 2, "FLT": 3}

    def set_ram_pressure(self, PPerT, PParT, scb, LZ, PHI, PressMode="SKD", iSm2=4, SavGolIters=11):
        """RAM pressures (nS,NR,NT) for the device front end of `pressure` (src/ModScbRun.f90:858-980): summed over the
        species%SCB species, extended radially, smoothed -- once per scb_run."""
        a, b = np.asfortranarray(PPerT, dtype=np.float64), np.asfortranarray(PParT, dtype=np.float64)
        nS, NR, NT = a.shape
        flags = np.ascontiguousarray(scb, dtype=np.int32)
        lz, ph = np.ascontiguousarray(LZ, dtype=np.float64), np.ascontiguousarray(PHI, dtype=np.float64)
        assert len(lz) >= NR + 1 and len(ph) >= NT
        _sck(self.L.rsg_scb_set_ram_pressure(self.h, nS, NR, NT, a.ctypes.data, b.ctypes.data, flags.ctypes.data, lz.ctypes.data,
                                             ph.ctypes.data, self.PRESS_MODES[PressMode], iSm2, SavGolIters))

    def get_ram_pressure(self):
        nX, nAz = C.c_int(), C.c_int()
        _sck(self.L.rsg_scb_get_ram_pressure(self.h, C.byref(nX), C.byref(nAz), None, None, None, None))
        r2, az = np.zeros(nX.value), np.zeros(nAz.value)
        per, par = np.zeros((nX.value, nAz.value), order="F"), np.zeros((nX.value, nAz.value), order="F")
        _sck(self.L.rsg_scb_get_ram_pressure(self.h, C.byref(nX), C.byref(nAz), r2.ctypes.data, az.ctypes.data, per.ctypes.data,
                                             par.ctypes.data))
        return r2, az, per, par

    def pressure_front(self, iLossCone=1, iReduceAnisotropy=0):
        """one `pressure` call entirely on the device; returns the normalised equatorial pressures (npsi, nzeta+1)"""
        pe = np.zeros((self.npsi, self.nzeta + 1), order="F")
        pa = np.zeros((self.npsi, self.nzeta + 1), order="F")
        _sck(self.L.rsg_scb_pressure_front(self.h, iLossCone, iReduceAnisotropy, pe.ctypes.data, pa.ctypes.data))
        return pe, pa

    def scb_run(self, pressure_fn, ordering=SOR_COLOR4, **kw):
        """scb_run (src/ModScbRun.f90:149-440) in one call, everything resident.  pressure_fn(xEq, yEq) ->
        (pperEq, pparEq), all (npsi, nzeta+1) Fortran-ordered: the 2-D front end of `pressure`."""
        p = ScbRunParams(InConAlpha=1e-6, InConPsi=1e-6, blendInitial=0.5, blendMin=0.01, blendMax=1.0, damp=0.9,
                         decreaseConvAlpha=0.5, decreaseConvPsi=0.5, nimax=5001, theChange=4, psiChange=0, numit=200,
                         MinSCBIterations=11, ordering=ordering, iLossCone=1, iReduceAnisotropy=0)
        for k, v in kw.items():
            if not hasattr(p, k):
                raise TypeError("unknown scb_run parameter " + k)
            setattr(p, k, v)
        shape = (self.npsi, self.nzeta + 1)
        err = []

        def cb(user, npsi, nzetap, xe, ye, pe, pa):
            try:
                n = npsi * nzetap
                x = np.ctypeslib.as_array(xe, shape=(n,)).reshape(shape, order="F")
                y = np.ctypeslib.as_array(ye, shape=(n,)).reshape(shape, order="F")
                a, b = pressure_fn(np.array(x, order="F"), np.array(y, order="F"))
                np.ctypeslib.as_array(pe, shape=(n,))[:] = np.asarray(a, dtype=np.float64).ravel(order="F")
                np.ctypeslib.as_array(pa, shape=(n,))[:] = np.asarray(b, dtype=np.float64).ravel(order="F")
                return 0
            except Exception as e:          # no exceptions across the C ABI
                err.append(e)
                return 1

        res = ScbRunResult()
        if pressure_fn is None:          # `pressure` front end on the device (set_ram_pressure before): no callback
            rc = self.L.rsg_scb_run(self.h, C.byref(p), C.cast(None, SCB_PRESSURE_FN), None, C.byref(res))
        else:
            rc = self.L.rsg_scb_run(self.h, C.byref(p), SCB_PRESSURE_FN(cb), None, C.byref(res))
        if err:
            raise err[0]
        _sck(rc)
        return {n: getattr(res, n) for n, _ in ScbRunResult._fields_}

    # ---- multi-GPU: iterateAlpha sharded along zeta (include/ramscb_gpu.h: rsg_scb_zsolve_*) ----
    def zsolve_begin(self, tol, k0, nk, nimax=5001, theChange=4, psiChange=0):
        _sck(self.L.rsg_scb_zsolve_begin(self.h, tol, nimax, theChange, psiChange, k0, nk))

    def zsolve_half(self, parity): _sck(self.L.rsg_scb_zsolve_half(self.h, parity))
    def zsolve_commit(self): _sck(self.L.rsg_scb_zsolve_commit(self.h))

    def zsolve_state_device(self):
        ptr, n = C.c_void_p(), C.c_longlong()
        _sck(self.L.rsg_scb_zsolve_state_device(self.h, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def zsolve_pending(self):
        n = C.c_int()
        _sck(self.L.rsg_scb_zsolve_pending(self.h, C.byref(n)))
        return n.value

    def field_device(self, name):
        ptr, n = C.c_void_p(), C.c_longlong()
        _sck(self.L.rsg_scb_field_device(self.h, name.encode(), C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def set_stream(self, stream_ptr):
        _sck(self.L.rsg_scb_set_stream(self.h, C.c_void_p(stream_ptr) if stream_ptr else None))

    def iterateAlpha(self, InConAlpha=1e-6, nimax=5001, theChange=4, psiChange=0, ordering=SOR_LEX):
        return self._iterate(self.L.rsg_scb_iterate_alpha, InConAlpha, nimax, theChange, psiChange, ordering, self.npsi)

    def iteratePsi(self, InConPsi=1e-6, nimax=5001, theChange=4, psiChange=0, ordering=SOR_LEX):
        return self._iterate(self.L.rsg_scb_iterate_psi, InConPsi, nimax, theChange, psiChange, ordering, self.nzeta)

    def Compute_convergence(self):
        a, b, c, f = C.c_double(), C.c_double(), C.c_double(), C.c_int()
        _sck(self.L.rsg_scb_convergence(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(f)))
        return {"normDiff": a.value, "normJxB": b.value, "normGradP": c.value, "SORFail": f.value}

    def derivs(self, f3):
        f3 = np.asfortranarray(f3, dtype=np.float64)
        out = [np.zeros(f3.shape, order="F") for _ in range(3)]
        _sck(self.L.rsg_scb_derivs(self.h, _p(f3), *[_p(o) for o in out]))
        return out

    def last_ms(self):
        return self.L.rsg_scb_last_ms(self.h)

    def use_cluster(self, on=True):
        _sck(self.L.rsg_scb_use_cluster(self.h, 1 if on else 0))

    def last_cluster(self):
        return self.L.rsg_scb_last_cluster(self.h)

    def launch_count(self):
        return self.L.rsg_scb_launch_count(self.h)

    def close(self):
        if self.h:
            self.L.rsg_scb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def hI_integrals(chiVal, mu, xRAM, yRAM, zRAM, bRAM, density, outsideMGNP, nThetaEquator, bnormal, HDens_cart=None, device=-1):
    """The integral block of computehI (src/ModRamScb.f90:372-410) for all RAM field lines in one launch:
    returns I_cart, H_cart, HDens_cart (nR,nT,nPa), bZEq_Cart (nR,nT) and the kernel's device time in ms.
    Arrays in the reference's shapes (Fortran order): xRAM .. density (nthe,nR,nT), outsideMGNP (nR,nT)."""
    L = _scb_lib()
    nthe, nR, nT = bRAM.shape
    nPa = len(mu)
    f = lambda a: np.asfortranarray(a, dtype=np.float64)
    chiVal, mu, xRAM, yRAM, zRAM, bRAM, density = (f(a) for a in (chiVal, mu, xRAM, yRAM, zRAM, bRAM, density))
    out = np.asfortranarray(outsideMGNP, dtype=np.int32)
    I = np.zeros((nR, nT, nPa), order="F")
    H = np.zeros((nR, nT, nPa), order="F")
    D = np.zeros((nR, nT, nPa), order="F") if HDens_cart is None else f(HDens_cart).copy(order="F")
    bz = np.zeros((nR, nT), order="F")
    ms = C.c_double(0.0)
    if device < 0:
        import torch
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    _sck(L.rsg_hI_integrals(device, nthe, nR, nT, nPa, int(nThetaEquator), float(bnormal), _p(chiVal), _p(mu), _p(xRAM), _p(yRAM),
                            _p(zRAM), _p(bRAM), _p(density), out.ctypes.data, _p(I), _p(H), _p(D), _p(bz), C.byref(ms)))
    return I, H, D, bz, ms.value


HI_RAM_NAMES = ("FNHS", "FNIS", "BOUNHS", "BOUNIS", "HDNS", "BNES")


def hI_tail(I_cart, H_cart, HDens_cart, bZEq_cart, ScaleAt, outsideMGNP, Lz, PA, PAbn, integral_smooth, DthI, ram, device=-1):
    """computehI after the integral block (src/ModRamScb.f90:413-637).  `ram`: dict with the previous FNHS, FNIS, BOUNHS,
    BOUNIS, HDNS (nR+1,nT,nPa) and BNES (nR+1,nT).  Returns a dict with the new RAM variables, dIdt, dHdt, dIbndt, dBdt,
    the four *_cart arrays as the reference leaves them, gslerr and the device time in ms."""
    L = _scb_lib()
    nR, nT, nPa = I_cart.shape
    f = lambda a: np.array(a, dtype=np.float64, order="F")
    out = {"I_cart": f(I_cart), "H_cart": f(H_cart), "HDens_cart": f(HDens_cart), "bZEq_cart": f(bZEq_cart)}
    for n in HI_RAM_NAMES:
        out[n] = f(ram[n])
    for n in ("dIdt", "dHdt", "dIbndt"):
        out[n] = np.zeros((nR + 1, nT, nPa), order="F")
    out["dBdt"] = np.zeros((nR + 1, nT), order="F")
    sa = np.ascontiguousarray(ScaleAt, dtype=np.int32)
    om = np.asfortranarray(outsideMGNP, dtype=np.int32)
    Lz, PA, PAbn = (np.ascontiguousarray(a, dtype=np.float64) for a in (Lz, PA, PAbn))
    err, ms = C.c_int(0), C.c_double(0.0)
    if device < 0:
        import torch
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    _sck(L.rsg_hI_tail(device, nR, nT, nPa, _p(out["I_cart"]), _p(out["H_cart"]), _p(out["HDens_cart"]), _p(out["bZEq_cart"]),
                       sa.ctypes.data, om.ctypes.data, _p(Lz), _p(PA), _p(PAbn), 1 if integral_smooth else 0, float(DthI),
                       *[_p(out[n]) for n in HI_RAM_NAMES], _p(out["dIdt"]), _p(out["dHdt"]), _p(out["dIbndt"]), _p(out["dBdt"]),
                       C.byref(err), C.byref(ms)))
    out["gslerr"], out["ms"] = err.value, ms.value
    return out


def hI_convert_lines(x, y, z, bf, psi, alfa, Lz, MLT, nThetaEquator, device=-1):
    """computehI's "Convert SCB field lines to RAM field lines" (src/ModRamScb.f90:252-300): returns xRAM, yRAM, zRAM,
    bRAM (nthe,nR,nT), outsideSCB (nR,nT) and the device time in ms.  SCB arrays (nthe,npsi,nzeta+1); Lz(nR+1), MLT(nT)."""
    L = _scb_lib()
    nthe, npsi, nz1 = x.shape
    f = lambda a: np.asfortranarray(a, dtype=np.float64)
    x, y, z, bf, psi, alfa = (f(a) for a in (x, y, z, bf, psi, alfa))
    Lz, MLT = (np.ascontiguousarray(a, dtype=np.float64) for a in (Lz, MLT))
    nR, nT = len(Lz) - 1, len(MLT)
    out = [np.zeros((nthe, nR, nT), order="F") for _ in range(4)]
    outside = np.zeros((nR, nT), dtype=np.int32, order="F")
    ms = C.c_double(0.0)
    if device < 0:
        import torch
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    _sck(L.rsg_hI_convert_lines(device, nthe, npsi, nz1 - 1, nR, nT, int(nThetaEquator), _p(x), _p(y), _p(z), _p(bf), _p(psi), _p(alfa),
                                _p(Lz), _p(MLT), *[_p(a) for a in out], outside.ctypes.data, C.byref(ms)))
    return (*out, outside, ms.value)


class HiGpu:
    """computehI resident on the device (rsg_hi, include/ramscb_gpu.h; src/ModRamScb.f90:249-637): the three blocks run
    back to back on one stream, the intermediate arrays never leave the device, HDens_cart and the RAM variables persist
    between calls like the reference's module arrays.  `scb` in convert()/computehI() is a dict with x, y, z, bf, psi, alfa
    (host arrays) or a ScbGpu (its device arrays are read in place)."""

    NAMES3 = ("xRAM", "yRAM", "zRAM", "bRAM", "density")
    NAMES_CART = ("I_cart", "H_cart", "HDens_cart")
    NAMES_RAM3 = ("FNHS", "FNIS", "BOUNHS", "BOUNIS", "HDNS", "dIdt", "dHdt", "dIbndt")
    NAMES_RAM2 = ("BNES", "dBdt")

    def __init__(self, nthe, npsi, nzeta, Lz, MLT, mu, PA, PAbn, chiVal, nThetaEquator, bnormal, device: int = -1):
        self.L = _scb_lib()
        c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        Lz, MLT, mu, PA, PAbn, chiVal = (c(a) for a in (Lz, MLT, mu, PA, PAbn, chiVal))
        self.nthe, self.npsi, self.nzeta = int(nthe), int(npsi), int(nzeta)
        self.nR, self.nT, self.nPa = len(Lz) - 1, len(MLT), len(mu)
        if device < 0:
            import torch
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        self.h = C.c_void_p()
        _sck(self.L.rsg_hi_create(C.byref(self.h), device, self.nthe, self.npsi, self.nzeta, self.nR, self.nT, self.nPa,
                                  int(nThetaEquator), float(bnormal), _p(chiVal), _p(mu), _p(Lz), _p(MLT), _p(PA), _p(PAbn)))

    def close(self):
        if self.h:
            self.L.rsg_hi_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_ram_fields(self, ram, HDens_cart=None):
        f = lambda a: np.asfortranarray(a, dtype=np.float64)
        a = [f(ram[n]) for n in HI_RAM_NAMES]
        hd = f(HDens_cart) if HDens_cart is not None else None
        _sck(self.L.rsg_hi_set_ram_fields(self.h, *[_p(x) for x in a], _p(hd) if hd is not None else None))

    def _scb_args(self, scb):
        if isinstance(scb, ScbGpu):
            return [None] * 6 + [scb.h], []
        f = lambda a: np.asfortranarray(a, dtype=np.float64)
        keep = [f(scb[n]) for n in ("x", "y", "z", "bf", "psi", "alfa")]
        return [_p(a) for a in keep] + [None], keep

    def convert(self, scb, want_outside=True):
        args, keep = self._scb_args(scb)
        out = np.zeros((self.nR, self.nT), dtype=np.int32, order="F") if want_outside else None
        n = C.c_int(0)
        _sck(self.L.rsg_hi_convert(self.h, *args, out.ctypes.data if out is not None else None, C.byref(n) if want_outside else None))
        return out, n.value

    def set_line(self, i, j, x, y, z, b):
        c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        x, y, z, b = c(x), c(y), c(z), c(b)
        _sck(self.L.rsg_hi_set_line(self.h, int(i), int(j), _p(x), _p(y), _p(z), _p(b)))

    def finish(self, DthI, integral_smooth=True, ScaleAt=None, outsideMGNP=None, density=None):
        sa = np.ascontiguousarray(ScaleAt, dtype=np.int32) if ScaleAt is not None else None
        om = np.asfortranarray(outsideMGNP, dtype=np.int32) if outsideMGNP is not None else None
        de = np.asfortranarray(density, dtype=np.float64) if density is not None else None
        err = C.c_int(0)
        _sck(self.L.rsg_hi_finish(self.h, sa.ctypes.data if sa is not None else None, om.ctypes.data if om is not None else None,
                                  _p(de) if de is not None else None, 1 if integral_smooth else 0, float(DthI), C.byref(err)))
        return err.value

    def computehI(self, scb, DthI, integral_smooth=True):
        args, keep = self._scb_args(scb)
        err = C.c_int(0)
        _sck(self.L.rsg_computehI(self.h, *args, 1 if integral_smooth else 0, float(DthI), C.byref(err)))
        return err.value

    def get(self, name):
        if name in self.NAMES3:
            out = np.zeros((self.nthe, self.nR, self.nT), order="F")
        elif name in self.NAMES_CART:
            out = np.zeros((self.nR, self.nT, self.nPa), order="F")
        elif name in self.NAMES_RAM3:
            out = np.zeros((self.nR + 1, self.nT, self.nPa), order="F")
        elif name in self.NAMES_RAM2:
            out = np.zeros((self.nR + 1, self.nT), order="F")
        else:
            out = np.zeros((self.nR, self.nT), order="F")          # psiRAM, bZEq_cart
        _sck(self.L.rsg_hi_get(self.h, name.encode(), _p(out)))
        return out

    def get_int(self, which):
        k = {"outsideSCB": 0, "outsideMGNP": 1, "ScaleAt": 2}[which]
        out = np.zeros(self.nT if k == 2 else (self.nR, self.nT), dtype=np.int32, order="F")
        _sck(self.L.rsg_hi_get_int(self.h, k, out.ctypes.data))
        return out

    def results(self):
        out = {n: self.get(n) for n in self.NAMES_RAM3 + self.NAMES_RAM2 + self.NAMES_CART + ("bZEq_cart", "xRAM", "yRAM", "zRAM", "bRAM")}
        for n in ("outsideSCB", "outsideMGNP", "ScaleAt"):
            out[n] = self.get_int(n)
        return out

    def push_to_ram(self, ram_gpu):
        """the new BNES .. dIbndt and outsideMGNP device-to-device into a RamGpu on the same device (rsg_ram_set_fields)"""
        ptrs = (C.c_void_p * 9)()
        om = C.c_void_p()
        _sck(self.L.rsg_hi_device_fields(self.h, C.cast(ptrs, C.c_void_p), C.cast(C.byref(om), C.c_void_p)))
        _ck(self.L.rsg_ram_set_fields_device(ram_gpu.h, C.cast(ptrs, C.c_void_p), om))

    def last_ms(self):
        return float(self.L.rsg_hi_last_ms(self.h))

    def launch_count(self):
        return int(self.L.rsg_hi_launch_count(self.h))


def computehI(scb, Lz, MLT, mu, PA, PAbn, ram, DthI, integral_smooth=True, density_fn=None, trace_fn=None, device=-1, _impl=None):
    """computehI (src/ModRamScb.f90:141-637, default branch) composed from the three device calls, with the reference's
    host-side steps in between: `scb` carries x, y, z, bf, psi, alfa (nthe,npsi,nzeta+1), chiVal, nThetaEquator, bnormal;
    `ram` the previous FNHS, FNIS, BOUNHS, BOUNIS, HDNS, BNES.  Lines outside the SCB domain set ScaleAt (:306-310) and go
    to `trace_fn(i, j, xo, yo)` -> (x, y, z, b) arrays of nthe nodes or None; without a tracer they are flagged
    outsideMGNP like the reference's 'SWMF' branch (:311-314).  `density_fn(distance)` is the bounce-averaged quantity
    (RAIRDEN polynomial in the reference, :365-371).  `_impl` (tests only) swaps the three calls for another
    implementation with the same signatures."""
    conv, integ, tail = _impl if _impl else (hI_convert_lines, hI_integrals, hI_tail)
    kw = {} if _impl else {"device": device}
    xR, yR, zR, bR, outsideSCB = conv(scb["x"], scb["y"], scb["z"], scb["bf"], scb["psi"], scb["alfa"], Lz, MLT, scb["nThetaEquator"], **kw)[:5]
    nR, nT = outsideSCB.shape
    ScaleAt = np.zeros(nT, dtype=np.int32)
    outsideMGNP = np.zeros((nR, nT), dtype=np.int32, order="F")
    for j in range(nT):
        for i in range(nR):
            if outsideSCB[i, j] == 1:
                if ScaleAt[j] == 0:
                    ScaleAt[j] = i + 1
                line = None
                if trace_fn is not None:
                    ang = MLT[j] * 2.0 * np.pi / 24.0 - np.pi
                    line = trace_fn(i, j, Lz[i + 1] * np.cos(ang), Lz[i + 1] * np.sin(ang))
                if line is None:
                    outsideMGNP[i, j] = 1
                else:
                    xR[:, i, j], yR[:, i, j], zR[:, i, j], bR[:, i, j] = line
    distance = np.sqrt(xR ** 2 + yR ** 2 + zR ** 2)
    density = np.asfortranarray(density_fn(distance) if density_fn else np.zeros_like(distance))
    I, H, D, bz = integ(scb["chiVal"], mu, xR, yR, zR, bR, density, outsideMGNP, scb["nThetaEquator"], scb["bnormal"], **kw)[:4]
    out = tail(I, H, D, bz, ScaleAt, outsideMGNP, Lz, PA, PAbn, integral_smooth, DthI, ram, **kw)
    out.update(xRAM=xR, yRAM=yR, zRAM=zR, bRAM=bR, outsideSCB=outsideSCB, outsideMGNP=outsideMGNP, ScaleAt=ScaleAt)
    return out
