"""Error convention of the reference: fatal conditions call terminate(code, ...), which never returns.
Each case runs in a subprocess and checks the exit code the library's fallback terminate() exits with.  No GPU needed:
all of these fire on the host before any device work."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PRELUDE = """
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, %r)
from kspace_neutrinos_b200 import capi
from tests import refs
L = capi.lib(); L.ksn_set_quiet(1)
om = refs.make_omnu(L); refs.set_background(L, om)
kk, dnu, dtot = refs.load_golden_state()
dcdm = refs.golden_delta_cdm(L, om, dnu, dtot)
tr = refs.load_transfer(L)
""" % ROOT

CASES = {
    # delta_tot_table.c:92-94: late start without a resume file
    2023: "d = refs.new_delta_tot(L, om, 300); L.delta_tot_init(C.byref(d), 300, refs.dptr(kk), refs.dptr(dcdm), C.byref(tr), 0.5)",
    # :95-97: before the transfer-function time
    2024: "d = refs.new_delta_tot(L, om, 300); L.delta_tot_init(C.byref(d), 300, refs.dptr(kk), refs.dptr(dcdm), C.byref(tr), 0.005)",
    # :98-100: more bins than allocated
    2011: "d = refs.new_delta_tot(L, om, 100); L.delta_tot_init(C.byref(d), 300, refs.dptr(kk), refs.dptr(dcdm), C.byref(tr), 0.01)",
    # :113-114: k beyond the CAMB table
    2: "k2 = kk * 1e6; d = refs.new_delta_tot(L, om, 300); L.delta_tot_init(C.byref(d), 300, refs.dptr(k2), refs.dptr(dcdm), C.byref(tr), 0.01)",
    # transfer_init.c:19-21: unreadable transfer file
    2019: "t = capi.TransferInitTable(); L.allocate_transfer_init_table(C.byref(t), 512000.0, 1.0, 1000.0, b'/nonexistent/file.dat')",
    # delta_tot_table.c:292-294: state file that starts at another scale factor
    2007: ("d = capi.DeltaTotTable(); L.allocate_delta_tot_table(C.byref(d), 300, 0.02, 1.0, refs.OMEGA0, C.byref(om), refs.UNIT_TIME, refs.UNIT_LENGTH, 0); "
           "L.read_all_nu_state(C.byref(d), os.path.join(refs.GOLDEN, 'delta_tot_nu.txt').encode())"),
}


@pytest.mark.parametrize("code", sorted(CASES))
def test_terminate_codes(code):
    r = subprocess.run([sys.executable, "-c", PRELUDE + textwrap.dedent(CASES[code])], capture_output=True, text=True, timeout=120)
    # exit codes are reported modulo 256
    assert r.returncode == code % 256, (r.returncode, r.stderr[-500:])
    assert r.stderr.strip(), "terminate() must say why"


def test_save_power_returns_minus_one_on_bad_directory():
    code = PRELUDE + "d = refs.new_delta_tot(L, om, 300); d.nk = 0; sys.exit(0 if L.save_nu_power(C.byref(d), 0.5, 3, b'/nonexistent/dir') == -1 else 1)"
    assert subprocess.run([sys.executable, "-c", code], capture_output=True, timeout=120).returncode == 0
