#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove the Blackwell paths (tcgen05 = UTCHMMA/UTCQMMA..., TMEM
# loads/stores = LDTM/STTM, TMA tensor / bulk copies = UTMALDG / UBLKCP, MUFU, packed FP32 = FFMA2...).
# usage: tools/sass_summary.sh [lib.so] > profiles/r2_sass_summary.txt
LIB=${1:-kernel_matrix_benchmarks_b200/libkmb_b200.so}
echo "# cuobjdump -sass $LIB  ($(cuobjdump -lelf "$LIB" | grep -c sm_100a) sm_100a cubins)"
echo "# kernel UTCHMMA LDTM STTM UTMALDG UBLKCP MUFU.EX2 MUFU.RSQ MUFU.SQRT FFMA2 FADD2 FMUL2 DFMA SHFL"
cuobjdump -sass "$LIB" | awk '
function flush() { if (name != "") printf "%s %d %d %d %d %d %d %d %d %d %d %d %d %d\n", name, c["UTCHMMA"], c["LDTM"], c["STTM"], c["UTMALDG"], c["UBLKCP"], c["MUFU.EX2"], c["MUFU.RSQ"], c["MUFU.SQRT"], c["FFMA2"], c["FADD2"], c["FMUL2"], c["DFMA"], c["SHFL"]; delete c }
/Function :/ { flush(); name = $3 }
/UTCHMMA/ { c["UTCHMMA"]++ } /LDTM/ { c["LDTM"]++ } /STTM/ { c["STTM"]++ } /UTMALDG/ { c["UTMALDG"]++ } /UBLKCP/ { c["UBLKCP"]++ }
/MUFU\.EX2/ { c["MUFU.EX2"]++ } /MUFU\.RSQ/ { c["MUFU.RSQ"]++ } /MUFU\.SQRT/ { c["MUFU.SQRT"]++ } /FFMA2/ { c["FFMA2"]++ } /FADD2/ { c["FADD2"]++ } /FMUL2/ { c["FMUL2"]++ } /DFMA/ { c["DFMA"]++ } /SHFL/ { c["SHFL"]++ }
END { flush() }' | while read -r name rest; do echo "$(echo "$name" | c++filt | cut -c1-110 | tr ' ' '_') $rest"; done
