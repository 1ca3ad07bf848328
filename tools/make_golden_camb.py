#!/usr/bin/env python3
"""Generate tests/golden/camb_linear_steps.npz from the reference's CAMB fixtures (run in the build
container, where /root/reference exists; the GPU box only sees the committed npz).

Restates load_camb_transfer (delta_tot_table_test.c:231-311): for each of the 99 scale factors
a = 0.01 ... 0.99 of test_reproduce_linear (:315-363) read the first NREAD=200 rows with k > 2 pi/512 of
ics_transfer_<a>.dat / ics_matterpow_<a>.dat and form
    delta_cdm = sqrt((T_cdm+b/T_tot)^2 P(k) 1000^3),  delta_nu = sqrt((T_nu/T_tot)^2 P(k) 1000^3),  keff = k/1000.
"""
import os
import shutil
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
NREAD, KMIN, SCALE = 200, 2 * np.pi / 512.0, 1000.0


def rows(path, ncol):
    out = []
    with open(path) as f:
        for line in f:
            if line.startswith("#"):
                continue
            tok = line.split()
            if len(tok) < ncol:
                break
            out.append([float(x) for x in tok[:ncol]])
    return np.array(out)


keffs, dcdm, dnu, avals = [], [], [], []
for i in range(99):
    a = 0.01 + i * 0.01
    tag = "%2g" % a                       # the reference builds file names with "%2g" (:347-348)
    t = rows(os.path.join(REF, "camb_linear", f"ics_transfer_{tag}.dat"), 8)
    m = rows(os.path.join(REF, "camb_linear", f"ics_matterpow_{tag}.dat"), 2)
    t = t[t[:, 0] > KMIN][:NREAD]
    m = m[m[:, 0] > KMIN][:NREAD]
    assert len(t) == NREAD and len(m) == NREAD
    k, T_nu, T_tot, T_cdm = t[:, 0], t[:, 5], t[:, 6], t[:, 7]
    pk = m[:, 1] * SCALE ** 3
    dcdm.append(np.sqrt((T_cdm / T_tot) ** 2 * pk))
    dnu.append(np.sqrt((T_nu / T_tot) ** 2 * pk))
    keffs.append(m[:, 0] / SCALE)
    avals.append(a)
np.savez_compressed(os.path.join(OUT, "camb_linear_steps.npz"), a=np.array(avals), keffs=np.array(keffs),
                    delta_cdm=np.array(dcdm), delta_nu_camb=np.array(dnu))
shutil.copyfile(os.path.join(REF, "camb_linear", "ics_transfer_0.01.dat"), os.path.join(OUT, "camb_ics_transfer_0.01.dat"))
print("wrote", OUT)
