"""Host-side mirror of the reference's per-step interface, on top of the C-ABI (capi.py).

The names follow the reference's domain: a *slab* is the contiguous range of the slowest grid index a
rank owns (`startslab`/`nslab` in powerspectrum.h:30, `slabstart_y`/`nslab_y` in interface_gadget.h:38);
a *grid* is the r2c-transformed density field, `nslab x PMGRID x (PMGRID/2+1)` complex values.
Everything that computes goes through libkspace_neutrinos_b200.so; nothing here falls back to a CPU path.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import capi

UNIT_LENGTH_KPC = 3.085678e21           # cm; Gadget's kpc/h unit (delta_tot_table_test.c:90)
UNIT_TIME = UNIT_LENGTH_KPC / 1e5       # s;  km/s velocity unit


@dataclass(frozen=True)
class Slab:
    """Planes [start, start+count) of the slowest grid index owned by one rank."""
    start: int
    count: int


def slab_partition(pmgrid: int, nranks: int) -> list[Slab]:
    """FFTW2-style slab decomposition: contiguous, as even as possible, first ranks get the extra planes."""
    if nranks < 1:
        raise ValueError("nranks must be >= 1")
    base, extra = divmod(pmgrid, nranks)
    out, start = [], 0
    for r in range(nranks):
        n = base + (1 if r < extra else 0)
        out.append(Slab(start, n))
        start += n
    return out


def modes_in_slab(pmgrid: int, slab: Slab) -> int:
    """Stored complex modes (the unit of BASELINE.json's metric) in a slab."""
    return slab.count * pmgrid * (pmgrid // 2 + 1)


@dataclass
class Cosmology:
    """The parameters the host N-body code hands over (README.txt:56-74, interface_common.h:12-30,66)."""
    transfer_file: str
    time_transfer: float = 0.01
    input_unit_length_in_cm: float = UNIT_LENGTH_KPC * 1e3
    mnu: tuple = (0.1, 0.1, 0.1)
    hybrid_neutrinos_on: int = 0
    vcrit: float = 500.0
    nu_crit_time: float = 0.333
    box_size: float = 512000.0
    unit_time_in_s: float = UNIT_TIME
    unit_length_in_cm: float = UNIT_LENGTH_KPC
    omega0: float = 0.2793
    hubble_param: float = 0.7
    tcmb0: float = 2.7255
    time_max: float = 1.0


class DeviceGrid:
    """A slab resident in HBM (allocated through the C-ABI, no torch involved)."""

    def __init__(self, pmgrid: int, slab: Slab, real_bytes: int = 8):
        self.lib = capi.lib()
        self.pmgrid, self.slab, self.real_bytes = pmgrid, slab, real_bytes
        self.nbytes = modes_in_slab(pmgrid, slab) * 2 * real_bytes
        self.ptr = C.c_void_p()
        capi.check(self.lib.ksn_device_malloc(C.byref(self.ptr), max(self.nbytes, 256)), "ksn_device_malloc")

    def fill_synthetic(self, seed: int = 20261017, slope: float = -1.0) -> None:
        capi.check(self.lib.ksn_fill_synthetic_grid(self.ptr, self.real_bytes, self.pmgrid, self.slab.start, self.slab.count, seed, slope),
                   "ksn_fill_synthetic_grid")

    def to_host(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            dt = np.float64 if self.real_bytes == 8 else np.float32
            out = np.empty((self.slab.count, self.pmgrid, self.pmgrid // 2 + 1, 2), dtype=dt)
        if self.nbytes:
            capi.check(self.lib.ksn_memcpy_d2h(out.ctypes.data_as(C.c_void_p), self.ptr, self.nbytes), "ksn_memcpy_d2h")
        return out

    def free(self) -> None:
        if self.ptr:
            self.lib.ksn_device_free(self.ptr)
            self.ptr = C.c_void_p()


class PinnedGrid:
    """A slab in page-locked host memory (what a PM code would hand to add_nu_power_to_rhogrid)."""

    def __init__(self, pmgrid: int, slab: Slab, real_bytes: int = 8):
        self.lib = capi.lib()
        self.pmgrid, self.slab, self.real_bytes = pmgrid, slab, real_bytes
        self.nbytes = modes_in_slab(pmgrid, slab) * 2 * real_bytes
        self.ptr = C.c_void_p()
        capi.check(self.lib.ksn_host_alloc_pinned(C.byref(self.ptr), max(self.nbytes, 256)), "ksn_host_alloc_pinned")

    def free(self) -> None:
        if self.ptr:
            self.lib.ksn_host_free_pinned(self.ptr)
            self.ptr = C.c_void_p()


class SlabFFT:
    """Device-resident slab FFT of the PM grid (ksn_fft_*): real space in x-slabs, rows padded to 2 (N/2+1) doubles, ->
    k space in y-slabs, FFTW's transposed order F[y][x][kz] -- the layout of gadget-2/0002 patch:116, i.e. what
    add_nu_power_to_rhogrid takes.  `gather`: for more than one rank, a callable that all-gathers a bytes object over the
    ranks in rank order (e.g. torch.distributed.all_gather_object); the peer-memory collective (init_p2p_from_torch)
    must be up on the same ranks."""

    def __init__(self, pmgrid: int, rank: int = 0, world: int = 1, gather=None):
        self.lib = capi.lib()
        L = self.lib
        self.pmgrid, self.rank, self.world = pmgrid, rank, world
        capi.check(L.ksn_fft_plan(pmgrid, world, rank), "ksn_fft_plan")
        v = [C.c_longlong() for _ in range(4)]
        rb, kb = C.c_size_t(), C.c_size_t()
        capi.check(L.ksn_fft_layout(*[C.byref(x) for x in v], C.byref(rb), C.byref(kb)), "ksn_fft_layout")
        self.xslab, self.yslab = Slab(v[0].value, v[1].value), Slab(v[2].value, v[3].value)
        self.real_bytes, self.kspace_bytes = rb.value, kb.value
        self.real, self.kspace = C.c_void_p(), C.c_void_p()
        capi.check(L.ksn_device_malloc(C.byref(self.real), max(self.real_bytes, 256)), "ksn_device_malloc")
        capi.check(L.ksn_device_malloc(C.byref(self.kspace), max(self.kspace_bytes, 256)), "ksn_device_malloc")
        handles = None
        if world > 1:
            if gather is None:
                raise ValueError("SlabFFT on more than one rank needs a `gather` callable")
            mine = (C.c_ubyte * 128)()
            capi.check(L.ksn_fft_export(self.kspace, self.real, mine), "ksn_fft_export")
            everyone = gather(bytes(mine))
            handles = (C.c_ubyte * (128 * world)).from_buffer_copy(b"".join(everyone))
        capi.check(L.ksn_fft_attach(self.kspace, self.real, handles), "ksn_fft_attach")

    def upload_real(self, rho: np.ndarray) -> None:
        """rho: this rank's x planes, [nslab_x][N][N] doubles (unpadded)."""
        n = self.pmgrid
        pad = np.zeros((self.xslab.count, n, 2 * (n // 2 + 1)))
        pad[..., :n] = rho
        if pad.nbytes:
            capi.check(self.lib.ksn_memcpy_h2d(self.real, pad.ctypes.data_as(C.c_void_p), pad.nbytes), "ksn_memcpy_h2d")

    def download_real(self) -> np.ndarray:
        n = self.pmgrid
        pad = np.empty((self.xslab.count, n, 2 * (n // 2 + 1)))
        if pad.nbytes:
            capi.check(self.lib.ksn_memcpy_d2h(pad.ctypes.data_as(C.c_void_p), self.real, pad.nbytes), "ksn_memcpy_d2h")
        return pad[..., :n].copy()

    def download_kspace(self) -> np.ndarray:
        n = self.pmgrid
        out = np.empty((self.yslab.count, n, n // 2 + 1, 2))
        if out.nbytes:
            capi.check(self.lib.ksn_memcpy_d2h(out.ctypes.data_as(C.c_void_p), self.kspace, out.nbytes), "ksn_memcpy_d2h")
        return out

    def forward(self) -> None:
        capi.check(self.lib.ksn_fft_forward(self.real, self.kspace), "ksn_fft_forward")

    def inverse(self) -> None:
        capi.check(self.lib.ksn_fft_inverse(self.kspace, self.real), "ksn_fft_inverse")

    def free(self) -> None:
        self.lib.ksn_fft_destroy()
        for p in (self.real, self.kspace):
            if p:
                self.lib.ksn_device_free(p)
        self.real, self.kspace = C.c_void_p(), C.c_void_p()


class KspaceNeutrinos:
    """The module-global integrator of the reference (one per process, like interface_common.c:17-26),
    driven in the host call order of SURVEY 3.1: parameters -> InitOmegaNu -> allocate_kspace_memory ->
    add_nu_power_to_rhogrid every PM step."""

    def __init__(self, cosmo: Cosmology, pmgrid: int, rank: int = 0, snapdir: str | None = None, quiet: bool = True):
        self.lib = capi.lib()
        self.cosmo, self.pmgrid, self.rank = cosmo, pmgrid, rank
        L = self.lib
        L.ksn_set_quiet(1 if quiet else 0)
        p = capi.kspace_params()
        p.KspaceTransferFunction = cosmo.transfer_file.encode()
        p.TimeTransfer = cosmo.time_transfer
        p.InputSpectrum_UnitLength_in_cm = cosmo.input_unit_length_in_cm
        for i in range(3):
            p.MNu[i] = cosmo.mnu[i]
        p.hybrid_neutrinos_on = cosmo.hybrid_neutrinos_on
        p.vcrit = cosmo.vcrit
        p.nu_crit_time = cosmo.nu_crit_time
        L.InitOmegaNu(cosmo.hubble_param, cosmo.tcmb0, 0)
        # stand-alone runs have no N-body code to supply hubble_function(): use the library's flat-LCDM fallback
        L.ksn_set_default_hubble(None, cosmo.omega0, cosmo.unit_time_in_s)
        dt = capi.global_delta_tot_table()
        dt.delta_tot_init_done = 0
        dt.ia = 0
        L.allocate_kspace_memory(pmgrid // 2, rank, cosmo.box_size, cosmo.unit_time_in_s, cosmo.unit_length_in_cm, cosmo.omega0,
                                 snapdir.encode() if snapdir else None, cosmo.time_max, 0)
        self.state = dt

    def add_nu_power_to_rhogrid(self, time: float, grid_ptr, slab: Slab, real_bytes: int = 8) -> None:
        """interface_gadget.h:38 -- grid_ptr may be a host or a device pointer."""
        fn = self.lib.add_nu_power_to_rhogrid_f64 if real_bytes == 8 else self.lib.add_nu_power_to_rhogrid_f32
        fn(time, self.cosmo.box_size, grid_ptr, self.pmgrid, slab.start, slab.count, 0)

    def add_nu_power_and_greens_to_rhogrid(self, time: float, grid_ptr, slab: Slab, asmth2: float, real_bytes: int = 8) -> None:
        """Extension: the same step with the PM Green's function (Gadget-2 pmforce_periodic) fused into the scaling pass."""
        fn = self.lib.add_nu_power_and_greens_to_rhogrid_f64 if real_bytes == 8 else self.lib.add_nu_power_and_greens_to_rhogrid_f32
        fn(time, self.cosmo.box_size, grid_ptr, self.pmgrid, slab.start, slab.count, asmth2, 0)

    def delta_nu_last(self) -> np.ndarray:
        return np.array([self.state.delta_nu_last[i] for i in range(self.state.nk)])

    def seed_history(self, rows: int = 98) -> None:
        """Install `rows` stored delta_tot rows at a = TimeTransfer*(1..rows) with delta_tot ~ a, through
        set_nu_state (interface_common.h:108).  Benchmarks use it to reach the long-history regime of
        BASELINE.json's configs without integrating 100 PM steps first."""
        st = self.state
        nk = st.nk
        a0 = self.cosmo.time_transfer
        sf = np.array([st.scalefact[0]] + [np.log(a0 * (i + 1)) for i in range(1, rows)])
        first = np.array([st.delta_tot[k][0] for k in range(nk)])
        dt = (first[:, None] * np.exp(sf - sf[0])[None, :]).copy()
        self.lib.set_nu_state(sf.ctypes.data_as(capi.c_double_p), dt.ctypes.data_as(capi.c_double_p), nk, rows, 0)


# ----------------------------------------------------------------------------- collectives
def init_nccl_from_torch(rank: int, world: int) -> None:
    """One process per GPU: create the library's own NCCL communicator (the per-step all-reduce of the bin
    sums runs on it, on the library's stream); the 128-byte unique id travels over torch.distributed."""
    import torch
    import torch.distributed as dist
    L = capi.lib()
    buf = (C.c_ubyte * 128)()
    if rank == 0:
        capi.check(L.ksn_comm_nccl_unique_id(buf), "ksn_comm_nccl_unique_id")
    obj = [bytes(buf)]
    dist.broadcast_object_list(obj, src=0)
    raw = (C.c_ubyte * 128).from_buffer_copy(obj[0])
    capi.check(L.ksn_comm_nccl_init(raw, world, rank), "ksn_comm_nccl_init")


def init_p2p_from_torch(rank: int, world: int) -> None:
    """One process per GPU of one box: the peer-memory backend (the cross-rank sum of the bin sums fused into the kernel
    that produces them, over NVLink).  Each rank exports its mailbox as a 64-byte CUDA-IPC handle; the handles are
    gathered over torch.distributed in rank order."""
    import torch.distributed as dist
    L = capi.lib()
    buf = (C.c_ubyte * 64)()
    capi.check(L.ksn_comm_p2p_export(buf), "ksn_comm_p2p_export")
    handles = [None] * world
    dist.all_gather_object(handles, bytes(buf))
    raw = (C.c_ubyte * (64 * world)).from_buffer_copy(b"".join(handles))
    capi.check(L.ksn_comm_p2p_init(raw, world, rank), "ksn_comm_p2p_init")
    dist.barrier()                      # nobody starts a round before every rank has mapped every mailbox


def init_comm_from_torch(rank: int, world: int, backend: str | None = None) -> str:
    """Pick the collective backend for a one-process-per-GPU run: KSN_COMM=p2p|nccl.  Default p2p (the cross-rank sum
    fused into the final-reduce kernel over peer memory); it is kept only if EVERY rank could map every mailbox and a
    trial all-reduce of a known vector came out right on every rank -- otherwise all ranks fall back to NCCL together.
    Returns the backend in use."""
    import sys
    import torch
    import torch.distributed as dist
    backend = backend or os.environ.get("KSN_COMM", "p2p")
    if backend == "p2p":
        L = capi.lib()

        def everyone(flag: bool) -> bool:          # collective: did it work on every rank?
            t = torch.tensor([1 if flag else 0], dtype=torch.int32, device="cuda" if torch.cuda.is_available() else "cpu")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            return int(t.item()) == 1

        why = ""
        handle = b""
        try:
            buf = (C.c_ubyte * 64)()
            capi.check(L.ksn_comm_p2p_export(buf), "ksn_comm_p2p_export")
            handle = bytes(buf)
        except RuntimeError as e:
            why = str(e)
        handles = [None] * world
        dist.all_gather_object(handles, handle)    # every rank takes part, whatever happened above
        ok = all(len(h) == 64 for h in handles)
        if ok:
            try:
                raw = (C.c_ubyte * (64 * world)).from_buffer_copy(b"".join(handles))
                capi.check(L.ksn_comm_p2p_init(raw, world, rank), "ksn_comm_p2p_init")
            except RuntimeError as e:
                ok, why = False, str(e)
        if everyone(ok):                           # (also the barrier: every mailbox is mapped before any round starts)
            try:
                v = np.full(1000, rank + 1.0)
                capi.check(L.ksn_comm_allreduce_host(v.ctypes.data_as(capi.c_double_p), v.size), "trial all-reduce")
                ok = bool(np.all(v == world * (world + 1) / 2))
                why = "" if ok else "trial all-reduce returned wrong sums"
            except RuntimeError as e:
                ok, why = False, str(e)
            if everyone(ok):
                return "p2p"
        if rank == 0:
            print(f"[ksn] peer-memory backend not usable on every rank ({why or 'another rank failed'}); using NCCL", file=sys.stderr)
    init_nccl_from_torch(rank, world)
    return "nccl"


_host_cb_keepalive = []


def init_host_allreduce_from_torch(rank: int, world: int) -> None:
    """Sum the bin partials with a torch.distributed all-reduce on host memory (gloo or any backend that
    takes CPU tensors).  This is the path an MPI host takes (MPI_Allreduce on its communicator)."""
    import torch
    import torch.distributed as dist
    L = capi.lib()

    def _cb(ptr, n, _user):
        arr = np.ctypeslib.as_array(ptr, shape=(n,))
        t = torch.from_numpy(arr)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return 0

    cb = capi.ALLREDUCE_FN(_cb)
    _host_cb_keepalive.append(cb)
    capi.check(L.ksn_comm_host_callback(cb, None, world, rank), "ksn_comm_host_callback")


def finish_powerspectrum(power_sum: np.ndarray, keff_sum: np.ndarray, count: np.ndarray, total_mass2: float):
    """The host tail of total_powerspectrum (powerspectrum.c:96-116) applied to all-reduced bin sums:
    returns (n_nonempty, power, count, keffs) compacted like the reference."""
    L = capi.lib()
    L.ksn_finish_powerspectrum.restype = C.c_int
    L.ksn_finish_powerspectrum.argtypes = [C.c_int, C.c_double, capi.c_double_p, capi.c_longlong_p, capi.c_double_p]
    p = np.array(power_sum, dtype=np.float64)
    k = np.array(keff_sum, dtype=np.float64)
    c = np.array(count, dtype=np.int64)
    n = L.ksn_finish_powerspectrum(len(p), float(total_mass2), p.ctypes.data_as(capi.c_double_p),
                                   c.ctypes.data_as(capi.c_longlong_p), k.ctypes.data_as(capi.c_double_p))
    return n, p, c, k
