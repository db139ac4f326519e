#!/bin/bash
# Builds variants of the library with different compile-time choices of the pair kernel and times them side by side
# on the GPU (bench_tools/pair_ab.py): how the constant-sign loop's form, Newton-reciprocal masks, unroll depth and the
# cost-model constants were chosen (profiles/r2c_pair_ab_variants.jsonl).  ptxas's schedule decides as much as the
# instruction counts, so every choice is an A/B on the device.
#
#   bash bench_tools/pair_variants.sh build  name1 "-DARVAE_NR_MASK_TP=0x10" name2 "-DARVAE_NR_MASK_TP=0x01 -DARVAE_CONST_OUTER_UNROLL=4" ...
#   bash bench_tools/pair_variants.sh run    [pair_ab.py environment, e.g. AB_WORKLOAD=c2_dsprites_b4096 AB_BATCH=65536]
#
# Macros (csrc/reg_sorted.cu): ARVAE_NR_MASK (16 bits, per-pair loop), ARVAE_NR_MASK_TP (8 bits, shared-reciprocal loop),
# ARVAE_CONST_OUTER_UNROLL, ARVAE_COST_GENERAL1, ARVAE_COST_TIE1.  (The forms that lost -- r = E_j / (E_i + E_j), both
# quotients of a quad, sums without staged T and P, the six-instruction Newton step -- are in the history of that file.)
# The variant libraries go to bench_tools/_exp/ (git-ignored; they travel to the GPU box with the snapshot).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REPO="$(dirname "$HERE")"
OUT="$HERE/_exp"
mkdir -p "$OUT"
cmd="$1"; shift || true
if [ "$cmd" = build ]; then
    while [ $# -ge 2 ]; do
        name="$1"; flags="$2"; shift 2
        ( cd "$REPO" && ARVAE_LIB_OUT="$OUT/lib_$name.so" ARVAE_OBJ_DIR="/tmp/arvae_obj_$name" ARVAE_NVCC_EXTRA="$flags" \
              python -m arvae_b200.build --force > "$OUT/$name.log" 2>&1 && echo "built $name ($flags)" ) &
    done
    wait
elif [ "$cmd" = run ]; then
    cd "$REPO" && env "$@" python bench_tools/pair_ab.py default "$OUT"/lib_*.so
else
    sed -n 2,14p "$0"
fi
