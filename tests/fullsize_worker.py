"""TEST WORKER (launched by tests/test_fullsize_parity_gpu.py through torch.distributed.run, one process per plane range,
all of them on cuda:0): the PRODUCT's total_powerspectrum and add_nu_power_to_rhogrid on a few planes of a full-width grid
(PMGRID 1024 / 2048 / 4096), device-resident, the bin sums reduced over the host call-back collective (gloo) -- the path an
MPI host takes.  Writes the input bytes for the reference (oracle/_ref/ref_slabs) and its own outputs in ref_slabs' format.

  fullsize_worker.py N hybrid m0 m1 m2 IN OUT R start_0 n_0 ... T a_0 ...
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch.distributed as dist
    from kspace_neutrinos_b200 import capi, host
    from tests import refs

    a = sys.argv[1:]
    n, hybrid = int(a[0]), int(a[1])
    masses = tuple(float(x) for x in a[2:5])
    in_path, out_path, R = a[5], a[6], int(a[7])
    slabs = [(int(a[8 + 2 * r]), int(a[9 + 2 * r])) for r in range(R)]
    T = int(a[8 + 2 * R])
    times = [float(x) for x in a[9 + 2 * R:9 + 2 * R + T]]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    rb = int(os.environ.get("KSN_TEST_REAL_BYTES", "8"))            # 4: float grids (the *_f32 entries)
    sfx = "_f64" if rb == 8 else "_f32"
    assert world == R
    L = capi.lib()
    capi.check(L.ksn_init(0), "ksn_init")
    dist.init_process_group(backend="gloo")
    if world > 1:
        host.init_host_allreduce_from_torch(rank, world)
    start, cnt = slabs[rank]
    slab = host.Slab(start, cnt)
    plane = n * (n // 2 + 1)
    off = sum(c for _, c in slabs[:rank]) * plane * 2 * rb
    grid = host.DeviceGrid(n, slab, rb)
    grid.fill_synthetic()
    g0 = grid.to_host()
    # the same bytes for the reference
    if rank == 0:
        with open(in_path, "wb") as f:
            f.truncate(sum(c for _, c in slabs) * plane * 2 * rb)
    dist.barrier()
    with open(in_path, "r+b") as f:
        f.seek(off)
        f.write(g0.tobytes())
    nb = n // 2
    spectra = []
    k1_names = []
    for sweep in range(2):           # first sweep of a geometry: the three-sum kernel; second: the tile kernel + cached geometry
        nret, P, Cn, K = refs.total_powerspectrum(L, g0, nb, startslab=start, nslab=cnt, fn="total_powerspectrum" + sfx, pointer=grid.ptr)
        spectra.append((nret, P[:nret].copy(), K[:nret].copy(), Cn[:nret].astype(np.float64)))
        k1_names.append(L.ksn_last_k1_kernel().decode())
    sim = host.KspaceNeutrinos(host.Cosmology(transfer_file=refs.default_transfer_file(), mnu=masses, hybrid_neutrinos_on=hybrid), n, rank=rank)
    dnus = []
    for t in times:
        sim.add_nu_power_to_rhogrid(t, grid.ptr, slab, rb)
        dnus.append(sim.delta_nu_last().copy())
    k3_name = L.ksn_last_k3_kernel().decode()
    k1_step = L.ksn_last_k1_kernel().decode()
    got = grid.to_host()
    if rank == 0:
        for idx, (nret, P, K, Cn) in enumerate(spectra):
            with open(out_path + (".first" if idx == 0 else ""), "wb") as f:
                f.write(np.array([n, nret, sim.state.nk, sim.state.ia, T], dtype=np.int32).tobytes())
                f.write(P.tobytes()); f.write(K.tobytes()); f.write(Cn.tobytes())
                for d in dnus:
                    f.write(np.array([float(len(d))]).tobytes()); f.write(d.tobytes())
        with open(out_path + ".grid", "wb") as f:
            f.truncate(sum(c for _, c in slabs) * plane * 2 * rb)
        with open(out_path + ".json", "w") as f:
            json.dump({"k1_first": k1_names[0], "k1_cached": k1_names[1], "k1_step": k1_step, "k3": k3_name}, f)
    dist.barrier()
    with open(out_path + ".grid", "r+b") as f:
        f.seek(off)
        f.write(got.tobytes())
    grid.free()
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank}/{world} ok")


if __name__ == "__main__":
    main()
