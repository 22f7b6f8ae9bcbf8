#!/bin/bash
# SASS evidence that the hot path is tcgen05 / TMEM / TMA (B200_PROFILING.md mnemonics), from the shipped library
SO=${1:-posepipeline_b200/libposeengine.so}
echo "# SASS mnemonic counts of $SO ($(date -u +%F), nvcc $(nvcc --version | grep release | sed 's/.*release //'))"
cuobjdump -sass $SO > /tmp/sass.txt
for m in UTCHMMA UTCQMMA "LDTM" "STTM" "UTMALDG" "UTMASTG" "UTCBAR" "SYNCS" "HMMA" "FFMA"; do
  printf "%-10s %8d\n" "$m" "$(grep -c "$m" /tmp/sass.txt)"
done
echo
echo "# per kernel (UTCHMMA / LDTM / UTMALDG / UTMASTG), conv_tc_kernel instantiations and the other kernels of the path"
awk '/Function :/ {name=$3} /UTCHMMA/ {a[name]++} /LDTM/ {b[name]++} /UTMALDG/ {c[name]++} /UTMASTG/ {d[name]++} /FFMA/ {f[name]++} END {for (n in a) printf "UTCHMMA=%-3d LDTM=%-3d UTMALDG=%-3d UTMASTG=%-3d FFMA=%-4d %s\n", a[n], b[n], c[n], d[n], f[n], n}' /tmp/sass.txt | c++filt | sed 's/(CUtensorMap_st.*//' | sort -k6 | head -100
