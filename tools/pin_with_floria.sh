#!/bin/bash
# pin_with_floria.sh — pins this repository's oracle (and through it the CUDA path) to the REAL floria.
#
# Needs what this image does not have: cargo + the floria crate's dependencies (crates.io or a vendor directory).  Run it on
# any machine with a Rust toolchain; bench.py and the tests never call it.  What it does:
#   1. builds tools/pin_harness (a 60-line Rust program over the unmodified floria library crate);
#   2. writes the pinning inputs as H-PoP fragment files (the format floria itself reads, file_reader.rs:37-109):
#        a) the fragments this repository's reader extracts from floria's own tests/test_long.bam + tests/test.vcf
#           (tests/golden/config0_long_frags.npz), b) a synthetic contig (configs[1] shape at 1/10 scale);
#   3. runs the harness at log level Debug, which makes floria dump every block's best partition to
#      local_parts/<j>-<l>-<snp_lo>-<ploidy>.haplosets (graph_processing.rs:289-300);
#   4. writes the same files from the oracle's fb_block_results with floria_b200/writers.py and diffs the two trees.
# A clean diff pins read selection, the ploidy loop, beam search, optimize_clustering and the stopping rule at once.
# Known, declared sources of difference: hashbrown iteration order at exact ties -> run the oracle with
# order_model=1 (FB_ORDER_MODEL=1) for the comparison.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
FLORIA_SRC="${FLORIA_SRC:-/root/reference}"
OUT="${1:-$ROOT/gpurun_out/pin}"
command -v cargo >/dev/null || { echo "pin_with_floria.sh: cargo not found (parity stays unpinned, DESIGN.md §5)"; exit 3; }
mkdir -p "$OUT"
# the harness depends on the floria crate by path
sed "s#path = \"../../../floria\"#path = \"$FLORIA_SRC\"#" "$HERE/pin_harness/Cargo.toml" > "$OUT/Cargo.toml"
mkdir -p "$OUT/src" && cp "$HERE/pin_harness/src/main.rs" "$OUT/src/main.rs"
(cd "$OUT" && cargo build --release)
python "$HERE/pin_inputs.py" "$OUT"            # writes <case>.hpop, <case>.snps, and ours/<case>/local_parts/*
rc=0
for case in config0 synth; do
  read -r eps mp bl < "$OUT/$case.params"
  "$OUT/target/release/floria-pin-harness" "$OUT/$case.hpop" "$OUT/$case.snps" "$OUT/floria/$case" "$eps" "$mp" "$bl"
  if diff -r "$OUT/floria/$case/local_parts" "$OUT/ours/$case/local_parts" > "$OUT/$case.diff"; then
    echo "PINNED: $case — $(ls "$OUT/ours/$case/local_parts" | wc -l) block partitions identical to floria"
  else
    echo "DIFFERENT: $case — see $OUT/$case.diff"; rc=1
  fi
done
exit $rc
