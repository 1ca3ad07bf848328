"""ctypes binding of libkspace_neutrinos_b200.so (the C-ABI declared in include/*.h).

This is the reference-side binding a maintainer would write: struct layouts and prototypes
follow include/kspace_neutrinos.h (the reference API) and include/ksn_b200.h (device layer).
No compute happens in Python; without the shared library, or without a CUDA device, the
compute entry points fail loudly -- there is no CPU fallback in this package (the CPU oracle
under oracle/ is test infrastructure and is never imported from here).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libkspace_neutrinos_b200.so")
NUSPECIES = 3

c_double_p = C.POINTER(C.c_double)
c_longlong_p = C.POINTER(C.c_longlong)


class RhoNuSingle(C.Structure):           # struct _rho_nu_single
    _fields_ = [("loga", c_double_p), ("rhonu", c_double_p), ("interp", C.c_void_p), ("acc", C.c_void_p), ("mnu", C.c_double)]


class HybridNu(C.Structure):              # struct _hybrid_nu
    _fields_ = [("enabled", C.c_int), ("nufrac_low", C.c_double * NUSPECIES), ("nu_crit_time", C.c_double), ("vcrit", C.c_double)]


class OmegaNu(C.Structure):               # struct _omega_nu
    _fields_ = [("RhoNuTab", C.POINTER(RhoNuSingle) * NUSPECIES), ("nu_degeneracies", C.c_int * NUSPECIES),
                ("rhocrit", C.c_double), ("kBtnu", C.c_double), ("tcmb0", C.c_double), ("hybnu", HybridNu)]


class TransferInitTable(C.Structure):     # struct _transfer_init_table
    _fields_ = [("NPowerTable", C.c_int), ("logk", c_double_p), ("T_nu", c_double_p)]


class DeltaPow(C.Structure):              # struct _delta_pow
    _fields_ = [("logkk", c_double_p), ("delta_ratio", c_double_p), ("spline", C.c_void_p), ("acc", C.c_void_p),
                ("nbins", C.c_int), ("norm", C.c_double)]


class DeltaTotTable(C.Structure):         # struct _delta_tot_table
    _fields_ = [("nk", C.c_int), ("nk_allocated", C.c_int), ("namax", C.c_int), ("ia", C.c_int), ("ThisTask", C.c_int),
                ("delta_nu_prefac", C.c_double), ("delta_tot_init_done", C.c_int), ("debug", C.c_int),
                ("delta_tot", C.POINTER(c_double_p)), ("scalefact", c_double_p), ("delta_nu_init", c_double_p),
                ("delta_nu_last", c_double_p), ("wavenum", c_double_p), ("omnu", C.POINTER(OmegaNu)),
                ("Omeganonu", C.c_double), ("light", C.c_double), ("TimeTransfer", C.c_double)]


class KspaceParams(C.Structure):          # struct __kspace_params
    _fields_ = [("KspaceTransferFunction", C.c_char * 500), ("TimeTransfer", C.c_double),
                ("InputSpectrum_UnitLength_in_cm", C.c_double), ("MNu", C.c_double * NUSPECIES),
                ("hybrid_neutrinos_on", C.c_int), ("vcrit", C.c_double), ("nu_crit_time", C.c_double)]


class DeltaNuArgs(C.Structure):           # struct ksn_delta_nu_args
    _fields_ = [("nk", C.c_int), ("Na", C.c_int), ("namax", C.c_int), ("nspecies", C.c_int),
                ("a", C.c_double), ("TimeTransfer", C.c_double), ("light", C.c_double), ("delta_nu_prefac", C.c_double),
                ("deriv_prefac", C.c_double), ("mnubykT", C.c_double * 3), ("qc", C.c_double * 3), ("nufrac_low0", C.c_double),
                ("relerr", C.c_double * 3), ("integrate", C.c_int * 3),
                ("scalefact", c_double_p), ("delta_tot", c_double_p), ("wavenum", c_double_p), ("delta_nu_init", c_double_p)]


class Timing(C.Structure):                # struct ksn_timing
    _fields_ = [("k1_ms", C.c_float), ("k1_reduce_ms", C.c_float), ("comm_ms", C.c_float), ("k2_ms", C.c_float),
                ("k3_ms", C.c_float), ("h2d_ms", C.c_float), ("d2h_ms", C.c_float), ("launches", C.c_ulonglong)]


HUBBLE_FN = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, c_double_p, C.c_size_t, C.c_void_p)

# name -> (restype, argtypes); every symbol include/*.h declares
PROTOTYPES = {
    # ---- ksn_b200.h
    "ksn_init": (C.c_int, [C.c_int]),
    "ksn_shutdown": (None, []),
    "ksn_last_error": (C.c_char_p, []),
    "ksn_device_available": (C.c_int, []),
    "ksn_device_count": (C.c_int, []),
    "ksn_device": (C.c_int, []),
    "ksn_comm_single": (C.c_int, []),
    "ksn_comm_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "ksn_comm_nccl_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "ksn_comm_p2p_export": (C.c_int, [C.c_void_p]),
    "ksn_comm_p2p_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "ksn_comm_host_callback": (C.c_int, [ALLREDUCE_FN, C.c_void_p, C.c_int, C.c_int]),
    "ksn_comm_allreduce_host": (C.c_int, [c_double_p, C.c_size_t]),
    "ksn_comm_rank": (C.c_int, []),
    "ksn_comm_size": (C.c_int, []),
    "ksn_device_malloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "ksn_device_free": (C.c_int, [C.c_void_p]),
    "ksn_host_alloc_pinned": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "ksn_host_free_pinned": (C.c_int, [C.c_void_p]),
    "ksn_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "ksn_host_unregister": (C.c_int, [C.c_void_p]),
    "ksn_memcpy_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "ksn_memcpy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "ksn_memcpy_d2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "ksn_device_synchronize": (C.c_int, []),
    "ksn_pointer_is_device": (C.c_int, [C.c_void_p]),
    "ksn_fill_synthetic_grid": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_longlong, C.c_ulonglong, C.c_double]),
    "ksn_powerspectrum_sums": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_longlong,
                                         C.POINTER(C.c_uint), c_double_p, c_double_p, c_double_p, c_longlong_p, c_double_p]),
    "ksn_scale_modes": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_longlong, C.c_double,
                                  c_double_p, c_double_p, C.c_int, C.c_double]),
    "ksn_scale_modes_greens": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_longlong, C.c_double,
                                        c_double_p, c_double_p, C.c_int, C.c_double, c_double_p, C.c_double]),
    "ksn_step_staged": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_longlong,
                                  C.POINTER(C.c_uint), c_double_p, C.c_double, C.c_void_p, C.c_void_p]),
    "ksn_step_staged_greens": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_longlong,
                                         C.POINTER(C.c_uint), c_double_p, C.c_double, C.c_void_p, C.c_void_p, C.c_double]),
    "ksn_delta_nu_integrate": (C.c_int, [C.POINTER(DeltaNuArgs), c_double_p, C.POINTER(C.c_ulonglong)]),
    "ksn_delta_nu_prefetch": (C.c_int, [C.c_double, C.c_double, C.c_double, c_double_p, C.c_int, C.c_int]),
    "ksn_last_k2_prefetch_used": (C.c_int, []),
    "ksn_k1_tile_plan": (C.c_int, [C.c_int, C.c_int, C.c_size_t, C.c_int] + [C.POINTER(C.c_int)] * 5),
    "ksn_k1_tile_plan_ex": (C.c_int, [C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int] + [C.POINTER(C.c_int)] * 5),
    "ksn_last_k1_kernel": (C.c_char_p, []),
    "ksn_last_k3_kernel": (C.c_char_p, []),
    "ksn_k3_table_plan": (C.c_int, [C.c_int, C.c_double, c_double_p, c_double_p, C.c_int, C.c_double,
                                    C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_uint), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ksn_k3_table_hash": (C.c_int, [C.c_int, C.c_double, c_double_p, c_double_p, C.c_int, C.c_double, C.c_int, C.POINTER(C.c_ulonglong)]),
    "ksn_last_k3_table": (C.c_int, [C.POINTER(c_double_p), C.POINTER(c_double_p), C.POINTER(C.c_int), c_double_p, c_double_p]),
    "ksn_last_k2_evals": (C.c_ulonglong, []),
    "ksn_last_k2_max_passes": (C.c_uint, []),
    "ksn_k2_spec_width": (C.c_int, []),
    "ksn_last_k2_max_trips": (C.c_uint, []),
    "ksn_set_background": (C.c_int, [HUBBLE_FN, C.c_void_p, C.c_double, C.c_double, C.c_int]),
    "ksn_background_info": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ksn_fslength_device": (C.c_int, [c_double_p, C.c_int, C.c_double, C.c_double, c_double_p]),
    "ksn_background_loaded": (C.c_int, []),
    "ksn_fft_plan": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "ksn_fft_layout": (C.c_int, [c_longlong_p, c_longlong_p, c_longlong_p, c_longlong_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "ksn_fft_export": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "ksn_fft_attach": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "ksn_fft_forward": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ksn_fft_inverse": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ksn_fft_timing": (C.c_int, [C.POINTER(C.c_float)]),
    "ksn_fft_destroy": (None, []),
    "ksn_timing_enable": (C.c_int, [C.c_int]),
    "ksn_timing_reset": (C.c_int, []),
    "ksn_timing_get": (C.c_int, [C.POINTER(Timing)]),
    "ksn_stream": (C.c_void_p, []),
    # ---- kspace_neutrinos.h: Omega_nu
    "rho_nu_init": (None, [C.POINTER(RhoNuSingle), C.c_double, C.c_double, C.c_double, C.c_double]),
    "rho_nu": (C.c_double, [C.POINTER(RhoNuSingle), C.c_double, C.c_double]),
    "init_hybrid_nu": (None, [C.POINTER(HybridNu), c_double_p, C.c_double, C.c_double, C.c_double, C.c_double]),
    "particle_nu_fraction": (C.c_double, [C.POINTER(HybridNu), C.c_double, C.c_int]),
    "nufrac_low": (C.c_double, [C.c_double]),
    "rho_nu_int": (C.c_double, [C.c_double, c_double_p]),       # params = {a*mnu, kT} (omega_nu_single.c:89)
    "get_rho_nu_conversion": (C.c_double, []),
    "init_omega_nu": (None, [C.POINTER(OmegaNu), c_double_p, C.c_double, C.c_double, C.c_double]),
    "get_omega_nu": (C.c_double, [C.POINTER(OmegaNu), C.c_double]),
    "get_omega_nu_nopart": (C.c_double, [C.POINTER(OmegaNu), C.c_double]),
    "get_omegag": (C.c_double, [C.POINTER(OmegaNu), C.c_double]),
    "omega_nu_single": (C.c_double, [C.POINTER(OmegaNu), C.c_double, C.c_int]),
    # transfer
    "allocate_transfer_init_table": (None, [C.POINTER(TransferInitTable), C.c_double, C.c_double, C.c_double, C.c_char_p]),
    "free_transfer_init_table": (None, [C.POINTER(TransferInitTable)]),
    # delta_pow
    "init_delta_pow": (None, [C.POINTER(DeltaPow), c_double_p, c_double_p, C.c_int, C.c_double]),
    "get_dnudcdm_powerspec": (C.c_double, [C.POINTER(DeltaPow), C.c_double]),
    "free_d_pow": (None, [C.POINTER(DeltaPow)]),
    # delta_tot_table
    "allocate_delta_tot_table": (None, [C.POINTER(DeltaTotTable), C.c_int, C.c_double, C.c_double, C.c_double,
                                        C.POINTER(OmegaNu), C.c_double, C.c_double, C.c_int]),
    "free_delta_tot_table": (None, [C.POINTER(DeltaTotTable)]),
    "delta_tot_init": (None, [C.POINTER(DeltaTotTable), C.c_int, c_double_p, c_double_p, C.POINTER(TransferInitTable), C.c_double]),
    "update_delta_tot": (None, [C.POINTER(DeltaTotTable), C.c_double, c_double_p, c_double_p, C.c_int]),
    "get_delta_nu_update": (None, [C.POINTER(DeltaTotTable), C.c_double, C.c_int, c_double_p, c_double_p, c_double_p,
                                   C.POINTER(TransferInitTable)]),
    "get_delta_nu": (None, [C.POINTER(DeltaTotTable), C.c_double, c_double_p, c_double_p, C.c_double]),
    "get_delta_nu_combined": (None, [C.POINTER(DeltaTotTable), C.c_double, c_double_p, c_double_p]),
    "save_delta_tot": (None, [C.POINTER(DeltaTotTable), C.c_int, C.c_char_p]),
    "save_all_nu_state": (None, [C.POINTER(DeltaTotTable), C.c_char_p]),
    "save_nu_power": (C.c_int, [C.POINTER(DeltaTotTable), C.c_double, C.c_int, C.c_char_p]),
    "read_all_nu_state": (None, [C.POINTER(DeltaTotTable), C.c_char_p]),
    "specialJ": (C.c_double, [C.c_double, C.c_double, C.c_double]),
    "fslength": (C.c_double, [C.c_double, C.c_double, C.c_double]),
    "get_delta_tot": (C.c_double, [C.c_double] * 6),
    # interface_common
    "OmegaNu": (C.c_double, [C.c_double]),
    "OmegaNu_nopart": (C.c_double, [C.c_double]),
    "InitOmegaNu": (None, [C.c_double, C.c_double, C.c_int]),
    "allocate_kspace_memory": (None, [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_char_p, C.c_double, C.c_int]),
    "compute_neutrino_power_from_cdm": (DeltaPow, [C.c_double, c_double_p, c_double_p, C.POINTER(C.c_long), C.c_int, C.c_int]),
    "save_nu_state": (None, [C.c_char_p]),
    "get_nu_state": (None, [C.POINTER(c_double_p), C.POINTER(c_double_p), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "set_nu_state": (None, [c_double_p, c_double_p, C.c_size_t, C.c_size_t, C.c_int]),
    "save_neutrino_power": (C.c_int, [C.c_double, C.c_int, C.c_char_p]),
    "particle_nu_active": (C.c_int, [C.c_double]),
    # interface_gadget / powerspectrum
    "set_kspace_vars": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int]),
    "save_total_power": (C.c_int, [C.c_double, C.c_int, C.c_char_p]),
}
for _name in ("total_powerspectrum", "total_powerspectrum_f64", "total_powerspectrum_f32"):
    PROTOTYPES[_name] = (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, c_double_p, c_longlong_p, c_double_p, C.c_int])
for _name in ("add_nu_power_to_rhogrid", "add_nu_power_to_rhogrid_f64", "add_nu_power_to_rhogrid_f32",
              "compute_total_power_spectrum", "compute_total_power_spectrum_f64", "compute_total_power_spectrum_f32"):
    PROTOTYPES[_name] = (None, [C.c_double, C.c_double, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int])
for _name in ("add_nu_power_and_greens_to_rhogrid_f64", "add_nu_power_and_greens_to_rhogrid_f32"):
    PROTOTYPES[_name] = (None, [C.c_double, C.c_double, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int])
# helpers exported for stand-alone use (src/ksn_host.h)
PROTOTYPES["ksn_set_default_hubble"] = (None, [C.POINTER(OmegaNu), C.c_double, C.c_double])
PROTOTYPES["ksn_set_quiet"] = (None, [C.c_int])
PROTOTYPES["hubble_function"] = (C.c_double, [C.c_double])
PROTOTYPES["ksn_global_omnu"] = (C.POINTER(OmegaNu), [])

_lib = None


def build(force: bool = False) -> str:
    """Compile the shared library in-tree (gcc + nvcc for sm_100a)."""
    if force:
        subprocess.run(["make", "-C", PKG_DIR, "clean"], check=True, capture_output=True)
    r = subprocess.run(["make", "-C", PKG_DIR, "-j8"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libkspace_neutrinos_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    return LIB_PATH


def lib() -> C.CDLL:
    """Load (once) and return the shared library with all prototypes attached."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback for this package)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)     # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


class KsnError(RuntimeError):
    pass


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise KsnError(f"{what or 'ksn call'} failed ({rc}): {lib().ksn_last_error().decode()}")


def kspace_params() -> KspaceParams:
    return KspaceParams.in_dll(lib(), "kspace_params")


def global_delta_tot_table() -> DeltaTotTable:
    return DeltaTotTable.in_dll(lib(), "delta_tot_table")
