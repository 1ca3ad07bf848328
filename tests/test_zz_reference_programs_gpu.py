"""GPU part of tests/test_reference_programs.py (named to sort last): the reference's own cmocka programs that reach
the kernels -- powerspectrum_test.c (K1 on the 4^3 known-answer grid, double and float builds) and
delta_tot_table_test.c (K2 behind get_delta_nu_update, save/resume, fslength, specialJ, the 99-step CAMB linear-theory
run) -- linked to the product's libkspace_neutrinos_b200.so (oracle/_ref/dropin_*_test; prebuilt, /root/reference is
not read at run time; their data files come from oracle/_ref/fixtures)."""
import pytest

from tests.test_reference_programs import dropin, fixtures_dir, n_run_failed, run

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,ntests", [("powerspectrum", 1), ("powerspectrum_single", 1), ("delta_tot_table", 7),
                                         ("omega_nu_single", 9), ("transfer_init", 1), ("delta_pow", 2)])
def test_reference_program_passes_against_the_product_on_the_gpu(gpu, name, ntests, tmp_path):
    rc, out = run(dropin(name), fixtures_dir(tmp_path))
    assert rc == 0, out[-3000:]
    assert n_run_failed(out) == (ntests, 0), out[-3000:]
