#!/bin/bash
P=tests/dev/determinism_probe2.py
python $P 2>&1 | tail -1
HI_TC_SPLITS=1 python $P 2>&1 | tail -1
HI_TC_SPLITS=2 python $P 2>&1 | tail -1
HI_TC_SPLITS=7 python $P 2>&1 | tail -1
HI_PAIR_STATIC=1 python $P 2>&1 | tail -1
HI_PAIR_CTAS=1 python $P 2>&1 | tail -1
SEQS="[(300, 4000)]" python $P 2>&1 | tail -1
SEQS="[(36, 4000)]" python $P 2>&1 | tail -1
SEQS="[(36, 4000)]" HI_TC_SPLITS=2 python $P 2>&1 | tail -1
PATH_ID=2 python $P 2>&1 | tail -1
PATH_ID=1 python $P 2>&1 | tail -1
HI_B200_LIB=hydrainfer_b200/lib/libhi_b200_r1pair.so HI_TC_SPLITS=2 python $P 2>&1 | tail -1
