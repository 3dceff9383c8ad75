#!/bin/bash
# SASS evidence that the tensor-core kernels issue tcgen05 / TMA instructions (B200_PROFILING.md): per kernel, the count of each Blackwell
# mnemonic in the sm_100a cubin of the in-tree library.  No GPU needed.  Usage: bash scripts/sass_census.sh > profiles/rNN_sass_census.txt
LIB=${1:-onnxstream_b200/csrc/libonnxstream_b200.so}
echo "# cuobjdump -sass $LIB | mnemonic census (UTCHMMA = tcgen05.mma kind::f16/bf16, UTCIMMA = kind::i8, UTMALDG / UTMASTG = TMA load / store,"
echo "# UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st, UTCATOMSWS = TMEM alloc / dealloc, UCGABAR = cluster barrier, SYNCS = mbarrier ops)"
cuobjdump -sass "$LIB" 2>/dev/null | awk '
/Function :/ { fn=$3 }
{ for (i=1;i<=NF;i++) if ($i ~ /^(UTCHMMA|UTCIMMA|UTCQMMA|UTMALDG|UTMASTG|UTCBAR|UTCATOMSWS|LDTM|STTM|SYNCS|UCGABAR_ARV|UCGABAR_WAIT|ELECT)/) { split($i,b,"."); c[fn" "b[1]]++ } }
END { for (k in c) print k, c[k] }' | sort | while read fn m n; do echo "$(echo "$fn" | c++filt | sed -E 's/^void //; s/\(CUtensorMap.*//; s/\(anonymous namespace\):://g') $m $n"; done
