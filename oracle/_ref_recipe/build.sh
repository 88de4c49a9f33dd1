#!/bin/sh
# Build the reference's own tra_adv_fct into a stand-alone driver (see README.md).  Never copies reference sources.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
SRC=${NEMO_SRC:-/root/reference}
FC=${FC:-gfortran}
OUT=$HERE/../_ref
if ! command -v "$FC" >/dev/null 2>&1; then echo "_ref_recipe: no Fortran compiler ($FC) on this machine -- nothing built: no compiled-binary pin (the source-text pin is tests/test_cpu_reference_exec.py)"; exit 0; fi
if [ ! -f "$SRC/src/OCE/TRA/traadv_fct.F90" ]; then echo "_ref_recipe: no NEMO source tree at $SRC -- nothing built"; exit 0; fi
mkdir -p "$OUT/build" && cd "$OUT/build"
O=$SRC/src/OCE
FFLAGS="-cpp -Dkey_nosignedzero -fdefault-real-8 -O3 -funroll-all-loops -fcray-pointer -ffree-line-length-none -I$O -I$O/LBC -I$O/TRA -I$O/DOM"
for f in $O/par_kind.F90 $O/par_oce.F90 $SRC/ext/IOIPSL/src/nc4interface.F90 $O/IOM/in_out_manager.F90 $O/DOM/dom_oce.F90 $O/LBC/lbcnfd.F90 \
         $O/LBC/lib_mpp.F90 $O/LBC/lbclnk.F90 $O/oce.F90 $O/lib_fortran.F90 $O/trc_oce.F90 $O/TRD/trdmxl_oce.F90 $O/TRD/trdvor_oce.F90 $O/TRD/trd_oce.F90 $O/DOM/phycst.F90 \
         $HERE/stubs.F90 $O/TRA/traadv_fct.F90 $HERE/driver.F90; do
    echo "$FC $(basename $f)"; $FC $FFLAGS -c "$f"
done
$FC -o "$OUT/fct_ref_driver" *.o
echo "_ref_recipe: built $OUT/fct_ref_driver"
