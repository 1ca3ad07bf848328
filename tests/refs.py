"""Test-side loaders for the two CPU checkers (TEST INFRASTRUCTURE, never imported by the package):

* ``ref_lib(double=True)``  -- oracle/_ref/libksref_{double,single}.so: the reference's own C sources
  compiled unmodified against the shims + mini-GSL (built by oracle/Makefile where /root/reference
  exists; the .so travels to the GPU box).  None when absent.
* ``oracle_lib()``          -- oracle/libksn_oracle.so: the plain-C restatement of the path, buildable
  anywhere from this repository alone.
Both expose the reference API, so the ctypes prototypes of the product binding are reused.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from kspace_neutrinos_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
GOLDEN = os.path.join(ROOT, "tests", "golden")

UNIT_LENGTH = 3.085678e21
UNIT_TIME = UNIT_LENGTH / 1e5
OMEGA0 = 0.2793
HUBBLE_PARAM = 0.7
T_CMB0 = 2.7255
BOX = 512000.0

_cache = {}


def _attach(handle, names):
    for name in names:
        res, args = capi.PROTOTYPES[name]
        fn = getattr(handle, name)
        fn.restype = res
        fn.argtypes = args
    return handle


REF_API = [n for n in capi.PROTOTYPES if not n.startswith("ksn_") and not n.endswith(("_f64", "_f32")) and n != "hubble_function"]


def ref_lib(double=True):
    key = "ref_d" if double else "ref_s"
    if key not in _cache:
        path = os.path.join(ORACLE_DIR, "_ref", "libksref_double.so" if double else "libksref_single.so")
        if not os.path.exists(path) and os.path.exists("/root/reference/powerspectrum.c"):
            subprocess.run(["make", "-C", ORACLE_DIR, "-s", "ref"], capture_output=True)
        if not os.path.exists(path):
            _cache[key] = None
        else:
            h = C.CDLL(path, mode=C.RTLD_LOCAL)
            _attach(h, REF_API)
            h.ksn_ref_set_background.restype = None
            h.ksn_ref_set_background.argtypes = [C.POINTER(capi.OmegaNu), C.c_double, C.c_double]
            h.hubble_function.restype = C.c_double
            h.hubble_function.argtypes = [C.c_double]
            _cache[key] = h
    return _cache[key]


def oracle_lib():
    if "oracle" not in _cache:
        path = os.path.join(ORACLE_DIR, "libksn_oracle.so")
        r = subprocess.run(["make", "-C", ORACLE_DIR, "-s", "libksn_oracle.so"], capture_output=True, text=True)
        if r.returncode != 0 or not os.path.exists(path):
            raise RuntimeError("cannot build the oracle restatement:\n" + r.stdout + r.stderr)
        h = C.CDLL(path, mode=C.RTLD_LOCAL)
        _cache["oracle"] = h
    return _cache["oracle"]


def default_transfer_file():
    """The CAMB transfer table of the reference's own tests (testdata/ics_transfer_99.dat, copied into tests/golden/)."""
    return os.path.join(GOLDEN, "ics_transfer_99.dat")


def dptr(a):
    return a.ctypes.data_as(capi.c_double_p)


def make_omnu(libh, masses=(0.15, 0.15, 0.15), a0=0.01):
    om = capi.OmegaNu()
    m = (C.c_double * 3)(*masses)
    libh.init_omega_nu(C.byref(om), m, a0, HUBBLE_PARAM, T_CMB0)
    return om


def random_grid(n, seed=1, dtype=np.float64, slope=-1.5):
    """Hermitian-consistent r2c grid of a real Gaussian field with a power-law spectrum,
    element (0,0,0) = n^3 (the "total mass" the reference normalises by)."""
    rng = np.random.default_rng(seed)
    real = rng.standard_normal((n, n, n))
    f = np.fft.rfftn(real)
    kx = np.fft.fftfreq(n, 1.0 / n)
    kz = np.arange(n // 2 + 1)
    k2 = kx[:, None, None] ** 2 + kx[None, :, None] ** 2 + kz[None, None, :] ** 2
    k2[0, 0, 0] = 1
    f *= k2 ** (slope / 2)
    f[0, 0, 0] = n ** 3
    out = np.empty((n, n, n // 2 + 1, 2), dtype=dtype)
    out[..., 0] = f.real
    out[..., 1] = f.imag
    return out


def kat_grid_4():
    """The 4^3 grid of powerspectrum_test.c:19-32 after the forward r2c FFT."""
    field = np.zeros((4, 4, 6))
    flat = field.reshape(-1)
    for i in range(32):
        flat[6 * (i // 4) + i % 4] = 1
    flat[0] = 2
    f = np.fft.rfftn(field[:, :, :4])
    out = np.empty((4, 4, 3, 2))
    out[..., 0] = f.real
    out[..., 1] = f.imag
    return out


class DeviceBuffer:
    """A slab resident in HBM, allocated through the C-ABI."""

    def __init__(self, libh, host_array):
        self.lib = libh
        self.nbytes = host_array.nbytes
        self.ptr = C.c_void_p()
        capi.check(libh.ksn_device_malloc(C.byref(self.ptr), max(self.nbytes, 16)), "ksn_device_malloc")
        self.upload(host_array)

    def upload(self, host_array):
        a = np.ascontiguousarray(host_array)
        if a.nbytes:
            capi.check(self.lib.ksn_memcpy_h2d(self.ptr, a.ctypes.data_as(C.c_void_p), a.nbytes), "h2d")

    def download(self, like):
        out = np.empty_like(like)
        if out.nbytes:
            capi.check(self.lib.ksn_memcpy_d2h(out.ctypes.data_as(C.c_void_p), self.ptr, out.nbytes), "d2h")
        return out

    def free(self):
        if self.ptr:
            self.lib.ksn_device_free(self.ptr)
            self.ptr = C.c_void_p()


def total_powerspectrum(libh, grid, nrbins, startslab=0, nslab=None, fn="total_powerspectrum", pointer=None):
    """Call <fn>(dims, grid, nrbins, startslab, nslab, power, count, keffs, comm) -> (nret, power, count, keffs)."""
    n = grid.shape[1]
    if nslab is None:
        nslab = grid.shape[0]
    power = np.zeros(nrbins)
    keffs = np.zeros(nrbins)
    count = np.zeros(nrbins, dtype=np.int64)
    p = pointer if pointer is not None else grid.ctypes.data_as(C.c_void_p)
    nret = getattr(libh, fn)(n, p, nrbins, startslab, nslab, dptr(power), count.ctypes.data_as(capi.c_longlong_p), dptr(keffs), 0)
    return nret, power, count, keffs


def k3_numpy(grid, startslab, box, logkk, ratio, norm):
    """numpy restatement of interface_gadget.c:163-188 + delta_pow.c:19-37 (oracle for K3 at small sizes)."""
    nslab, n = grid.shape[0], grid.shape[1]
    ky = np.arange(startslab, startslab + nslab)
    ky = np.where(ky > n // 2, ky - n, ky).astype(np.float64)
    kx = np.arange(n)
    kx = np.where(kx > n // 2, kx - n, kx).astype(np.float64)
    kz = np.arange(n // 2 + 1, dtype=np.float64)
    k2 = ky[:, None, None] ** 2 + kx[None, :, None] ** 2 + kz[None, None, :] ** 2
    out = grid.astype(grid.dtype, copy=True)
    live = k2 > 0
    x = np.log(np.sqrt(np.where(live, k2, 1.0)) * 2 * np.pi / box)
    x = np.clip(x, logkk[0], logkk[-1])
    idx = np.clip(np.searchsorted(logkk, x, side="right") - 1, 0, len(logkk) - 2)
    y = ratio[idx] + (x - logkk[idx]) / (logkk[idx + 1] - logkk[idx]) * (ratio[idx + 1] - ratio[idx])
    smth = np.where(live, 1 + norm * y, 1.0)
    out[..., 0] = (grid[..., 0].astype(np.float64) * smth).astype(grid.dtype)
    out[..., 1] = (grid[..., 1].astype(np.float64) * smth).astype(grid.dtype)
    return out


def greens_numpy(grid, startslab, asmth2):
    """numpy restatement of the Green's-function loop of Gadget-2's pmforce_periodic (pm_periodic.c, the loop that follows
    the hook of gadget-2/0002 patch:116-125; Gadget-2 itself is not part of the reference repository): for k2 > 0
    smth = -exp(-k2*asmth2)/k2, f = sinc(pi k/N) per axis, smth *= (1/(fx fy fz))^4; F(0,0,0) = 0."""
    nslab, n = grid.shape[0], grid.shape[1]
    def kvals(idx):
        return np.where(idx > n // 2, idx - n, idx).astype(np.float64)
    ky, kx, kz = kvals(np.arange(startslab, startslab + nslab)), kvals(np.arange(n)), np.arange(n // 2 + 1, dtype=np.float64)
    def sinc(k):
        f = np.pi * k / n
        return np.where(k != 0, np.sin(np.where(k != 0, f, 1.0)) / np.where(k != 0, f, 1.0), 1.0)
    k2 = ky[:, None, None] ** 2 + kx[None, :, None] ** 2 + kz[None, None, :] ** 2
    ff = 1.0 / (sinc(ky)[:, None, None] * sinc(kx)[None, :, None] * sinc(kz)[None, None, :])
    live = k2 > 0
    smth = np.where(live, -np.exp(-k2 * asmth2) / np.where(live, k2, 1.0) * ff * ff * ff * ff, 0.0)
    out = np.empty_like(grid)
    out[..., 0] = (grid[..., 0].astype(np.float64) * smth).astype(grid.dtype)
    out[..., 1] = (grid[..., 1].astype(np.float64) * smth).astype(grid.dtype)
    return out


def greens_gadget2(grid, startslab, asmth2):
    """The C restatement of GADGET-2.0.7's Green's-function loop (oracle/gadget2_greens.c) on a copy of `grid`."""
    out = np.ascontiguousarray(grid).copy()
    orc().orc_gadget2_greens(out.ctypes.data_as(C.c_void_p), 1 if grid.dtype == np.float64 else 0, grid.shape[1], startslab, grid.shape[0], asmth2)
    return out


# ----------------------------------------------------------------------------- integrator fixtures
def load_golden_state():
    """The arrays delta_tot_table_test.c:setup_delta_pow (:367-452) builds from testdata/:
    k bins, delta_nu (sqrt P_nu) and delta_cdm at a=0.3333 for 3 x 0.15 eV."""
    with open(os.path.join(GOLDEN, "powerspec_nu_004.txt")) as f:
        tok = f.read().split()
    nb = int(tok[1])
    vals = np.array(tok[2:2 + 2 * nb], dtype=np.float64).reshape(nb, 2)
    kk = vals[:, 0].copy()
    delta_nu = np.sqrt(vals[:, 1])
    with open(os.path.join(GOLDEN, "powerspec_cdm_004.txt")) as f:
        tok = f.read().split()
    assert int(tok[1]) == nb
    delta_tot = np.array(tok[2:2 + nb], dtype=np.float64)
    return kk, delta_nu, delta_tot


def golden_delta_cdm(libh, om, delta_nu, delta_tot):
    OmegaNua3 = libh.get_omega_nu(C.byref(om), 0.01) * 0.01 ** 3
    OmegaMa = OMEGA0 - libh.get_omega_nu(C.byref(om), 1.0) + OmegaNua3
    fnu = OmegaNua3 / OmegaMa
    return (delta_tot - fnu * delta_nu) / (1.0 - fnu)


def set_background(libh, om):
    """Install the flat-LCDM + neutrino + photon Hubble rate of the reference's tests on either library."""
    _cache[("background", id(libh))] = om      # the library keeps a POINTER to it: it must outlive the caller's locals
    if hasattr(libh, "ksn_ref_set_background"):
        libh.ksn_ref_set_background(C.byref(om), OMEGA0, UNIT_TIME)
    else:
        libh.ksn_set_default_hubble(C.byref(om), OMEGA0, UNIT_TIME)


def load_transfer(libh, path=None, box=BOX):
    t = capi.TransferInitTable()
    path = path or os.path.join(GOLDEN, "ics_transfer_99.dat")
    libh.allocate_transfer_init_table(C.byref(t), box, UNIT_LENGTH, UNIT_LENGTH * 1e3, path.encode())
    return t


def new_delta_tot(libh, om, nk, time_transfer=0.01, time_max=1.0):
    d = capi.DeltaTotTable()
    libh.allocate_delta_tot_table(C.byref(d), nk, time_transfer, time_max, OMEGA0, C.byref(om), UNIT_TIME, UNIT_LENGTH, 0)
    return d


def init_module(libh, n, masses=(0.15, 0.15, 0.15), hybrid=False, time_transfer=0.01, box=BOX, transfer=None):
    """kspace_params -> InitOmegaNu -> allocate_kspace_memory, the host call order of SURVEY 3.1."""
    p = capi.KspaceParams.in_dll(libh, "kspace_params")
    p.KspaceTransferFunction = (transfer or os.path.join(GOLDEN, "ics_transfer_99.dat")).encode()
    p.TimeTransfer = time_transfer
    p.InputSpectrum_UnitLength_in_cm = UNIT_LENGTH * 1e3
    for i in range(3):
        p.MNu[i] = masses[i]
    p.hybrid_neutrinos_on = 1 if hybrid else 0
    p.vcrit = 500.0
    p.nu_crit_time = 0.333
    libh.InitOmegaNu(HUBBLE_PARAM, T_CMB0, 0)
    # a private _omega_nu with the same parameters drives the test Hubble rate of either library
    om = make_omnu(libh, masses, time_transfer)
    set_background(libh, om)
    dt = capi.DeltaTotTable.in_dll(libh, "delta_tot_table")
    dt.delta_tot_init_done = 0
    dt.ia = 0
    # the module allocates delta_cdm_last once, sized for the first grid it sees (interface_gadget.c:85-86);
    # forget it so that a re-initialisation with a larger grid does not overrun it
    C.c_void_p.in_dll(libh, "delta_cdm_last").value = None
    libh.allocate_kspace_memory(n // 2, 0, box, UNIT_TIME, UNIT_LENGTH, OMEGA0, None, 1.0, 0)
    return om, dt


# ----------------------------------------------------------------------------- oracle restatement binding
class OrcSpecies(C.Structure):
    _fields_ = [("mnu", C.c_double), ("tabulated", C.c_int), ("loga", C.c_double * 200), ("rho", C.c_double * 200),
                ("spline", C.c_void_p), ("acc", C.c_void_p)]


class OrcCosmo(C.Structure):
    _fields_ = [("sp", OrcSpecies * 3), ("degeneracy", C.c_int * 3), ("rhocrit", C.c_double), ("kBtnu", C.c_double),
                ("tcmb0", C.c_double), ("hybrid_on", C.c_int), ("nufrac_low", C.c_double * 3), ("nu_crit_time", C.c_double),
                ("vcrit", C.c_double), ("Omega_nonu", C.c_double), ("OmegaLambda", C.c_double), ("Hubble_internal", C.c_double)]


class OrcDtot(C.Structure):
    _fields_ = [("nk", C.c_int), ("nk_allocated", C.c_int), ("namax", C.c_int), ("ia", C.c_int), ("init_done", C.c_int),
                ("delta_nu_prefac", C.c_double), ("Omeganonu", C.c_double), ("light", C.c_double), ("TimeTransfer", C.c_double),
                ("scalefact", capi.c_double_p), ("delta_tot", capi.c_double_p), ("delta_nu_init", capi.c_double_p),
                ("delta_nu_last", capi.c_double_p), ("wavenum", capi.c_double_p), ("cosmo", C.POINTER(OrcCosmo)),
                ("n_evals", C.c_ulonglong)]


class OrcModule(C.Structure):
    _fields_ = [("cosmo", OrcCosmo), ("dtot", OrcDtot), ("t_logk", capi.c_double_p), ("t_tnu", capi.c_double_p), ("nt", C.c_int),
                ("scratch", capi.c_double_p), ("last_prefac", C.c_double), ("last_nk", C.c_int)]


def orc():
    """The oracle restatement with prototypes attached."""
    h = oracle_lib()
    if getattr(h, "_ksn_ready", False):
        return h
    D, I, LL, V = C.c_double, C.c_int, C.c_longlong, C.c_void_p
    dp, llp = capi.c_double_p, capi.c_longlong_p
    cp, tp, mp = C.POINTER(OrcCosmo), C.POINTER(OrcDtot), C.POINTER(OrcModule)
    protos = {
        "orc_cosmo_init": (None, [cp, dp, D, D, D]),
        "orc_cosmo_hybrid": (None, [cp, dp, D, D]),
        "orc_cosmo_background": (None, [cp, D, D]),
        "orc_omega_nu": (D, [cp, D]), "orc_omega_nu_nopart": (D, [cp, D]), "orc_omega_nu_single": (D, [cp, D, I]),
        "orc_omegag": (D, [cp, D]), "orc_particle_nu_fraction": (D, [cp, D, I]), "orc_nufrac_low": (D, [D]),
        "orc_hubble": (D, [cp, D]),
        "orc_powerspectrum_sums": (None, [I, V, I, I, LL, LL, dp, dp, llp, dp]),
        "orc_powerspectrum_finish": (I, [I, D, dp, llp, dp]),
        "orc_total_powerspectrum": (I, [I, V, I, I, LL, LL, dp, llp, dp]),
        "orc_dnudcdm": (D, [dp, dp, I, D, D]),
        "orc_scale_modes": (None, [V, I, I, LL, LL, D, dp, dp, I, D]),
        "orc_gadget2_greens": (None, [V, I, I, LL, LL, D]),
        "orc_fill_synthetic_grid": (None, [dp, I, LL, LL, C.c_ulonglong, D]),
        "orc_dtot_alloc": (None, [tp, I, D, D, D, cp, D, D]),
        "orc_dtot_free": (None, [tp]),
        "orc_dtot_read": (I, [tp, C.c_char_p]),
        "orc_transfer_read": (I, [C.c_char_p, D, D, D, C.POINTER(dp), C.POINTER(dp)]),
        "orc_dtot_init": (None, [tp, I, dp, dp, dp, dp, I, D]),
        "orc_fslength": (D, [cp, D, D, D]),
        "orc_specialJ": (D, [D, D, D]),
        "orc_get_delta_nu": (None, [tp, D, dp, dp, D]),
        "orc_get_delta_nu_combined": (None, [tp, D, dp, dp]),
        "orc_update_delta_tot": (None, [tp, D, dp, dp, I]),
        "orc_get_delta_nu_update": (I, [tp, D, I, dp, dp, dp, dp, dp, I]),
        "orc_module_init": (I, [mp, I, dp, I, D, D, C.c_char_p, D, D, D, D, D, D, D, D, D]),
        "orc_add_nu_power_to_rhogrid": (I, [mp, D, D, V, I, I, LL, LL]),
    }
    for name, (res, args) in protos.items():
        fn = getattr(h, name)
        fn.restype = res
        fn.argtypes = args
    h._ksn_ready = True
    return h


def orc_cosmo(masses=(0.15, 0.15, 0.15), a0=0.01, hybrid=False):
    o = orc()
    c = OrcCosmo()
    m = (C.c_double * 3)(*masses)
    o.orc_cosmo_init(C.byref(c), m, a0, HUBBLE_PARAM, T_CMB0)
    if hybrid:
        o.orc_cosmo_hybrid(C.byref(c), m, 500.0, 0.333)
    o.orc_cosmo_background(C.byref(c), OMEGA0, UNIT_TIME)
    return c


def orc_module(n, masses=(0.15, 0.15, 0.15), hybrid=False, time_transfer=0.01, box=BOX, transfer=None):
    o = orc()
    m = OrcModule()
    mm = (C.c_double * 3)(*masses)
    path = (transfer or os.path.join(GOLDEN, "ics_transfer_99.dat")).encode()
    rc = o.orc_module_init(C.byref(m), n // 2, mm, 1 if hybrid else 0, 500.0, 0.333, path, time_transfer, box, UNIT_TIME, UNIT_LENGTH,
                           UNIT_LENGTH * 1e3, OMEGA0, HUBBLE_PARAM, T_CMB0, 1.0)
    assert rc == 0, rc
    return m
