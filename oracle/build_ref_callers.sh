#!/bin/bash
# Builds the reference's OWN callers of the HYPREDRV_* API, unmodified, from the sources where they
# lie under /root/reference, against this repo's include/ and libHYPREDRV.so.  They are callers (an
# example driver and a unit test), not an implementation of the path: the arithmetic they reach is
# this library's.  Outputs go to oracle/_ref/ only (git-ignored, shipped to the GPU box).
#   examples/src/C_laplacian/laplacian.c   -> oracle/_ref/laplacian
#   tests/test_setmatrix_from_csr.c        -> oracle/_ref/test_setmatrix_from_csr
# The unit test includes the reference's internal headers (internal/error.h, internal/linsys.h,
# tests/test_helpers.h): they are taken from /root/reference at build time; HYPRE.h, mpi.h and
# HYPREDRV.h resolve to this repo's include/ first.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${REFERENCE_ROOT:-/root/reference}"
[ -d "$REF" ] || { echo "no reference tree at $REF: keeping prebuilt oracle/_ref" ; exit 0; }
mkdir -p "$HERE/_ref"
CFLAGS="-O1 -std=gnu11 -w -I$ROOT/include -I$REF/include -I$REF/tests"
LIBS="-L$ROOT/hypredrive_b200/lib -lHYPREDRV -lm -Wl,-rpath,\$ORIGIN/../../hypredrive_b200/lib"
gcc $CFLAGS -o "$HERE/_ref/laplacian" "$REF/examples/src/C_laplacian/laplacian.c" $LIBS
gcc $CFLAGS -o "$HERE/_ref/test_setmatrix_from_csr" "$REF/tests/test_setmatrix_from_csr.c" $LIBS
echo "built oracle/_ref/laplacian oracle/_ref/test_setmatrix_from_csr"
