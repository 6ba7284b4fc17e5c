#!/bin/bash
# Development aid: builds libkzg.so variants with extra nvcc flags into scripts/tmp_libs/ (CPU side) ...
#   bash scripts/variant_sweep.sh build name1 "-DX=1" name2 "-DY=2" ...
# ... and times them on the GPU box (the .so files travel with the snapshot):
#   gpurun -- 'bash scripts/variant_sweep.sh run "--kind logs"'
set -e
cd "$(dirname "$0")/.."
mkdir -p scripts/tmp_libs
if [ "$1" = build ]; then
  shift
  while [ $# -gt 1 ]; do
    name=$1; flags=$2; shift 2
    KZG_NVCC_EXTRA="$flags" python -m kompressor_b200.build > /dev/null
    cp kompressor_b200/libkzg.so scripts/tmp_libs/libkzg_$name.so
    echo "built $name ($flags)"
  done
  python -m kompressor_b200.build > /dev/null       # the default library again
else
  shift || true
  for f in scripts/tmp_libs/libkzg_*.so; do
    echo "== $f"
    KZG_LIB_PATH=$PWD/$f python scripts/dev_decode_bench.py $1 --variants "KZG_EXEC=1" 2>&1 | grep -o '"parity": [a-z]*\|"ms_per_step": [0-9.]*\|"fse": [0-9.]*\|"huffman": [0-9.]*\|"execute": [0-9.]*' | paste -sd' '
  done
fi
