# usage: tools/r02_ab_test.sh OUT variantA variantB ...  -- A/B of pre-built variants (twice each), then the GPU tests on the in-tree build
out=$1; shift
AB_STREAMS=256 AB_SECONDS=30 tools/ab_variants.sh gpurun_out/$out "$@" "$@" | cut -c1-260
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "^FAILED|^ERROR|passed|failed|Error" | head -20
