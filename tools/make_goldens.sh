#!/bin/bash
# make_goldens.sh — pin the oracle against REAL reference output.  Needs a JDK (javac + java); the build image has none,
# so this is for a box that does (SURVEY.md §8c mitigation 3, DESIGN.md "parity unpinned").
#
#   tools/make_goldens.sh /path/to/kanzi            # the flanglet/kanzi checkout (56399d4, Kanzi 2.5.0)
#
# For every case below it writes the synthetic input (python -m kanzi_b200.synth, pure numpy, seeded), runs the reference's
# own CLI on it with one job, and records tests/golden/<case>.knz (when small) plus size and sha256 in
# tests/golden/manifest.json.  `pytest tests/test_golden.py` then compares the oracle's (and, with a GPU, the CUDA path's)
# .knz with these byte for byte; without a manifest the test is skipped and the oracle stays "unpinned".
set -euo pipefail
REF=${1:?usage: make_goldens.sh /path/to/kanzi-checkout}
HERE=$(cd "$(dirname "$0")/.." && pwd)
command -v javac >/dev/null || { echo "javac not found: run this on a machine with a JDK"; exit 2; }
WORK=$(mktemp -d)
trap 'rm -rf "$WORK"' EXIT
echo "compiling the reference (java/src/main/java) into $WORK/classes"
mkdir -p "$WORK/classes"
find "$REF/java/src/main/java" -name '*.java' > "$WORK/sources.txt"
javac -nowarn -d "$WORK/classes" @"$WORK/sources.txt"
mkdir -p "$HERE/tests/golden"
MAN="$HERE/tests/golden/manifest.json"
echo '{"reference": "flanglet/kanzi java CLI, -j 1", "cases": [' > "$MAN"
first=1
# name | generator | bytes | seed | -t | -e | -b (bytes)
while IFS='|' read -r name gen n seed tr en bs; do
  [ -z "$name" ] && continue
  in="$WORK/$name.bin"; out="$WORK/$name.knz"
  (cd "$HERE" && python -m kanzi_b200.synth "$gen" "$n" "$seed" "$in")
  java -cp "$WORK/classes" io.github.flanglet.kanzi.app.Kanzi -c -f -j 1 -i "$in" -o "$out" -t "$tr" -e "$en" -b "$bs" >/dev/null
  size=$(stat -c %s "$out"); sha=$(sha256sum "$out" | cut -d' ' -f1)
  keep=false
  if [ "$size" -le 262144 ]; then cp "$out" "$HERE/tests/golden/$name.knz"; keep=true; fi
  [ $first -eq 1 ] || echo ',' >> "$MAN"; first=0
  printf '  {"name": "%s", "generator": "%s", "bytes": %s, "seed": %s, "transforms": "%s", "entropy": "%s", "block": %s, "knz_bytes": %s, "sha256": "%s", "knz_file": %s}' \
    "$name" "$gen" "$n" "$seed" "$tr" "$en" "$bs" "$size" "$sha" "$keep" >> "$MAN"
  echo "  $name: $size bytes, sha256 $sha"
done <<'CASES'
cfg1_huffman_64k|ascii_markov|1048576|1|NONE|HUFFMAN|65536
text_lz_ans0|text|300000|11|LZ|ANS0|65536
exe_lzx_huffman|exe_like|400000|12|LZX|HUFFMAN|131072
records_rolz_ans0|records|300000|13|ROLZ|ANS0|262144
text_bwt_rank_zrlt_ans1|text|262144|14|BWT+RANK+ZRLT|ANS1|131072
text_bwt_srt_zrlt_fpaq|text|262144|15|BWT+SRT+ZRLT|FPAQ|131072
noise_lz_ans0|noise|200000|16|LZ|ANS0|65536
cfg2_first_blocks|silesia_like|16777216|2|LZ|ANS0|4194304
cfg3_first_block|enwik_like|8388608|3|BWT+RANK+ZRLT|ANS1|8388608
pasted_lzp_ans0|pasted|400000|17|LZP|ANS0|131072
pasted_lzp_zrlt_huffman|pasted|300000|18|LZP+ZRLT|HUFFMAN|65536
mixed_rlt_fpaq|mixed_entropy|300000|19|RLT|FPAQ|131072
records_rlt_ans0|records|300000|20|RLT|ANS0|65536
text_rolzx_none|text|300000|21|ROLZX|NONE|131072
exe_rolzx_ans0|exe_like|300000|22|ROLZX|ANS0|262144
text_none_range|text|200000|23|NONE|RANGE|65536
text_bwts_rank_zrlt_ans0|text|262144|24|BWTS+RANK+ZRLT|ANS0|131072
CASES
echo ']}' >> "$MAN"
echo "wrote $MAN; commit tests/golden/ and run: python -m pytest tests/test_golden.py -q"
