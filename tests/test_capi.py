"""The C-ABI shared library: loads, exports every symbol include/*.h declares, and refuses to compute
without a CUDA device (no CPU fallback).  No GPU needed."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from kspace_neutrinos_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
    names = set()
    for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", text):
        nm = m.group(1)
        if nm in ("defined", "sizeof", "pow", "__attribute__", "alias", "int", "double", "void", "char"):
            continue                                   # keywords in casts / function-pointer typedefs
        names.add(nm)
    return names


def test_every_declared_symbol_is_exported(ksn):
    decl = _declared_functions("ksn_b200.h") | _declared_functions("kspace_neutrinos.h")
    # typedef'd callback types and macro names are not symbols
    decl -= {"ksn_allreduce_fn", "ksn_between_fn", "ksn_hubble_fn", "mymalloc", "myfree"}
    missing = [n for n in sorted(decl) if not hasattr(ksn, n)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"
    assert len(decl) > 80
    # and the binding covers them
    unbound = [n for n in sorted(decl) if n not in capi.PROTOTYPES and n not in ("terminate", "message", "mymalloc_fullinfo", "myfree_fullinfo")]
    assert not unbound, f"no ctypes prototype for: {unbound}"


def test_forwarding_headers_exist():
    for h in ("interface_common.h", "interface_gadget.h", "powerspectrum.h", "delta_tot_table.h", "delta_pow.h",
              "omega_nu_single.h", "transfer_init.h", "kspace_neutrino_const.h", "gadget_defines.h"):
        assert "kspace_neutrinos.h" in open(os.path.join(ROOT, "include", h)).read()


_LAYOUT_PROBE = r"""
#include <stdio.h>
#include <stddef.h>
%s
int main(void) {
  printf("%%zu %%zu %%zu %%zu %%zu %%zu %%zu\n", sizeof(struct _delta_tot_table), sizeof(struct _delta_pow), sizeof(struct _transfer_init_table),
         sizeof(kspace_params), sizeof(struct _hybrid_nu), sizeof(struct _omega_nu), sizeof(struct _rho_nu_single));
  printf("%%zu %%zu %%zu %%zu %%zu\n", offsetof(struct _delta_tot_table, delta_tot), offsetof(struct _delta_tot_table, omnu),
         offsetof(struct _delta_tot_table, TimeTransfer), offsetof(struct _delta_pow, norm), offsetof(struct _omega_nu, hybnu));
  return 0; }
"""


def _probe(tmp_path, includes, incdirs, tag):
    import subprocess
    src = tmp_path / f"probe_{tag}.c"
    src.write_text(_LAYOUT_PROBE % includes)
    exe = tmp_path / f"probe_{tag}"
    cmd = ["gcc", "-DPERIODIC", "-DDOUBLEPRECISION_FFTW"] + [f"-I{d}" for d in incdirs] + [str(src), "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()


def test_struct_layouts_match_the_reference_abi(tmp_path):
    """ctypes structs == include/kspace_neutrinos.h == (where present) the reference's own headers."""
    ours = _probe(tmp_path, '#include "kspace_neutrinos.h"', [os.path.join(ROOT, "include")], "ours")
    ct = [C.sizeof(capi.DeltaTotTable), C.sizeof(capi.DeltaPow), C.sizeof(capi.TransferInitTable), C.sizeof(capi.KspaceParams),
          C.sizeof(capi.HybridNu), C.sizeof(capi.OmegaNu), C.sizeof(capi.RhoNuSingle),
          capi.DeltaTotTable.delta_tot.offset, capi.DeltaTotTable.omnu.offset, capi.DeltaTotTable.TimeTransfer.offset,
          capi.DeltaPow.norm.offset, capi.OmegaNu.hybnu.offset]
    assert [int(x) for x in ours] == ct
    if os.path.exists("/root/reference/interface_common.h"):
        ref = _probe(tmp_path, '#include "interface_common.h"\n#include "delta_tot_table.h"\n#include "delta_pow.h"',
                     ["/root/reference", os.path.join(ROOT, "oracle", "shim")], "ref")
        assert ref == ours


def test_c_header_compiles_as_a_host_would_include_it(tmp_path):
    """A PM code keeps `#include "interface_gadget.h"`; the typed macro binds to the double build."""
    src = tmp_path / "host.c"
    src.write_text('#define DOUBLEPRECISION_FFTW\n#include "interface_gadget.h"\n#include "powerspectrum.h"\n'
                   'int main(void){ fftw_complex *g = 0; double p[4], k[4]; long long c[4];\n'
                   ' if (0) { add_nu_power_to_rhogrid(0.5, 1.0, g, 4, 0, 4, MPI_COMM_WORLD);\n'
                   '          total_powerspectrum(4, g, 4, 0, 4, p, c, k, MPI_COMM_WORLD); }\n'
                   ' return sizeof(fftw_complex) == 16 ? 0 : 1; }\n')
    exe = tmp_path / "host"
    import subprocess
    r = subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L", capi.PKG_DIR,
                        "-lkspace_neutrinos_b200", "-lm", f"-Wl,-rpath,{capi.PKG_DIR}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert subprocess.run([str(exe)]).returncode == 0


@pytest.mark.skipif(capi.lib().ksn_device_available() == 1, reason="a GPU is present")
def test_no_cpu_fallback(ksn):
    """Without a device every compute entry fails loudly."""
    assert ksn.ksn_init(-1) == -1                                   # KSN_ENODEV
    assert b"no CPU path" in ksn.ksn_last_error()
    g = np.zeros((4, 4, 3, 2))
    thr = (C.c_uint * 15)()
    iw = (C.c_double * 3)(1, 1, 1)
    out = np.zeros(15)
    cnt = np.zeros(15, dtype=np.int64)
    m2 = C.c_double()
    rc = ksn.ksn_powerspectrum_sums(g.ctypes.data_as(C.c_void_p), 8, 4, 15, 0, 4, thr, iw, out.ctypes.data_as(capi.c_double_p),
                                    out.ctypes.data_as(capi.c_double_p), cnt.ctypes.data_as(capi.c_longlong_p), C.byref(m2))
    assert rc == -1
    lk = np.array([0.0, 1.0])
    assert ksn.ksn_scale_modes(g.ctypes.data_as(C.c_void_p), 8, 4, 0, 4, 1.0, lk.ctypes.data_as(capi.c_double_p),
                               lk.ctypes.data_as(capi.c_double_p), 2, 0.1) == -1
