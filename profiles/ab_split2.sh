#!/bin/bash
# fused vs worklist path for the plans that run fused by default
for p in fused split; do echo "== GPUMOTIF_PATH=$p"; GPUMOTIF_PATH=$p bash profiles/quickbench.sh p$p 1024 score.1 ire mp.ends efn descr.quad descr.trip; done
