#!/bin/bash
# SASS evidence that the GEMM is tcgen05 / TMEM / TMA (B200_PROFILING.md "What proves a Blackwell-native kernel"):
# counts of the tensor-core, tensor-memory and TMA mnemonics in the compiled objects.  Runs without a GPU.
#   scripts/sass_summary.sh > profiles/gemm_sass_summary.txt
set -e
cd "$(dirname "$0")/.."
python -m mtrl_b200.build > /dev/null
echo "# cuobjdump -sass of mtrl_b200/build/*.o (sm_100a), built from the sources at $(git rev-parse --short HEAD 2>/dev/null || echo '?')"
echo "# mnemonic counts per object; PTX -> SASS: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, tcgen05.commit -> UTCBAR,"
echo "# cp.async.bulk.tensor -> UTMALDG/UTMASTG, cp.reduce.async.bulk.tensor -> UTMAREDG; HMMA would be the legacy mma.sync path"
for o in mtrl_b200/build/*.o; do
  n=$(cuobjdump -sass "$o" 2>/dev/null | grep -cE "UTC[A-Z]*MMA|LDTM|STTM|UTMALDG|UTMASTG|UTMAREDG|UTCBAR|HMMA" || true)
  echo "== $(basename "$o"): $n matching instructions"
  cuobjdump -sass "$o" 2>/dev/null | grep -oE "UTC[A-Z]*MMA(\.[A-Z0-9]+)*|LDTM(\.[A-Za-z0-9]+)*|STTM(\.[A-Za-z0-9]+)*|UTMALDG(\.[A-Z0-9]+)*|UTMASTG(\.[A-Z0-9]+)*|UTMAREDG(\.[A-Z0-9]+)*|UTCBAR(\.[A-Z0-9]+)*|HMMA[.A-Z0-9]*" | sort | uniq -c || true
done
echo "# kernels in gemm_tcgen05.o:"
cuobjdump -sass mtrl_b200/build/gemm_tcgen05.o 2>/dev/null | grep -E "Function :" | sed 's/^\s*/  /'
