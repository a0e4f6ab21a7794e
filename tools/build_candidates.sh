#!/bin/bash
# Builds the candidate libraries that have been checked on the host and by their SASS but not yet
# timed on a GPU; `gpurun -- 'bash tools/gpu_quad_variants.sh'` then runs the
# parity tests and the headline bench on each of them.
set -e
cd "$(dirname "$0")/.."
bash tools/build_variant.sh scanrec -DQUASI_SCAN_RECORDS=1
bash tools/build_variant.sh sincostab -DQUASI_SINCOS_TABLE=1
bash tools/build_variant.sh scanrec_sincostab -DQUASI_SCAN_RECORDS=1 -DQUASI_SINCOS_TABLE=1
bash tools/build_variant.sh lookup -DQUASI_SINCOS_LOOKUP=1
bash tools/build_variant.sh scanrec_lookup -DQUASI_SCAN_RECORDS=1 -DQUASI_SINCOS_LOOKUP=1   # 60 B of spills in the demo-scene kernel
bash tools/build_variant.sh pdl -DQUASI_PDL=1
bash tools/build_variant.sh scanrec_sincostab_pdl -DQUASI_SCAN_RECORDS=1 -DQUASI_SINCOS_TABLE=1 -DQUASI_PDL=1
bash tools/build_variant.sh scanrec_lookup_pdl -DQUASI_SCAN_RECORDS=1 -DQUASI_SINCOS_LOOKUP=1 -DQUASI_PDL=1
ls -la quasi_b200/lib/variants/
