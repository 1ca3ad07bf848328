import sys; sys.path.insert(0,'.')
import ctypes as C, numpy as np, math
from tests import refs
from kspace_neutrinos_b200 import capi
L = capi.lib(); capi.check(L.ksn_init(-1)); L.ksn_set_quiet(1)
om = refs.make_omnu(L); refs.set_background(L, om)
hub = capi.HUBBLE_FN(lambda a, _u: L.hubble_function(a))
for n in (4096, 16384, 65536):
    capi.check(L.ksn_set_background(hub, None, math.log(0.01)-0.01, 0.01, n))
    lo = np.array([math.log(x) for x in (0.01,0.02,0.05,0.1,0.2,0.3,0.5,0.9,0.99)])
    out = np.zeros(len(lo))
    capi.check(L.ksn_fslength_device(refs.dptr(lo), len(lo), 0.0, 299792., refs.dptr(out)))
    host = np.array([L.fslength(x, 0.0, 299792.) for x in lo])
    print(n, (out/host-1))
