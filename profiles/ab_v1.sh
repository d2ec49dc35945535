#!/bin/bash
# profiles/ab_v1.sh -- sieve_word_r variants: base / pair-bitset words hoisted per 32 span offsets, noinline (v1) / inlined (v1i)
bash profiles/ab_libs.sh 1024 "trna ire score.1 pk1 qu+tr" lib_vb.so lib_v1.so lib_v1i.so
