#!/bin/bash
# TEST INFRASTRUCTURE: run the reference's five cmocka programs (reference Makefile:18-21,
# "make test" = BASELINE.json configs[0]) built by oracle/Makefile against the shims and
# mini-GSL.  test_save_resume writes into testdata/ (delta_tot_table_test.c:113), so the
# fixtures are copied to a scratch directory first; /root/reference is never written.
set -u
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
make -s -C "$HERE" ref >/dev/null 2>&1 || { echo "build failed"; exit 2; }
S=$(mktemp -d)
cp -r "$REF/testdata" "$S/testdata"
ln -s "$REF/camb_linear" "$S/camb_linear"
rc=0
cd "$S"
for t in omega_nu_single transfer_init powerspectrum delta_pow delta_tot_table; do
    echo "=== ${t}_test"
    "$HERE/_ref/${t}_test" || rc=1
done
rm -rf "$S"
exit $rc
