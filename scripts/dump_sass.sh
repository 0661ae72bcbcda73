#!/bin/bash
# SASS listings of the hot kernels from the in-tree library (cuobjdump; no GPU needed) -> profiles/
SO=bldfm_b200/libbldfm_b200.so
cuobjdump -sass $SO > /tmp/all.sass
ext() { awk -v pat="$1" '/Function : /{f = ($0 ~ pat)} f' /tmp/all.sass | grep -v "^\s*/\* 0x" > "$2"; }
ext "k_marchILi0ELb0" profiles/sass_march_exact.txt
ext "k_marchILi1ELb0" profiles/sass_march_fma.txt
ext "k_marchILi2ELb0" profiles/sass_march_sweep.txt
ext "k_fft24IdLi0ELi6" profiles/sass_fft24_passX_c128_q64.txt
ext "k_fft24IdLi1ELi6" profiles/sass_fft24_passY_c128_q64.txt
ext "k_fft48IdLi0ELi5" profiles/sass_fft48_passX_c128_q32.txt
ext "k_fft48IdLi1ELi5" profiles/sass_fft48_passY_c128_q32.txt
ext "k_fft24pIdLi0ELi6" profiles/sass_fft24p_passX_c128_q64.txt
ext "k_fft_hIdLi0E" profiles/sass_fft_h_passX_c128.txt
ext "k_fft_hIdLi1E" profiles/sass_fft_h_passY_c128.txt
wc -l profiles/sass_*.txt
grep -c "LDG.E.ENL2.256" profiles/sass_fft24_passY_c128_q64.txt profiles/sass_fft48_passY_c128_q32.txt
grep -c "UBLKCP\|SYNCS" profiles/sass_fft24p_passX_c128_q64.txt profiles/sass_march_sweep.txt
