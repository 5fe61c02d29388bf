#!/bin/bash
# same-box A/B of two builds of the library (ab/libmpc_old.so, ab/libmpc_new.so): the table-law probe with each, twice
set -e
SO=mpconstellation_b200/csrc/libmpc_b200.so
for round in 1 2; do
  for v in ${VARIANTS:-old new}; do
    cp ab/libmpc_$v.so $SO; touch $SO
    echo "== $v (round $round)"
    python scripts/r02_probe_sequence_quick.py 2>&1 | tail -2
  done
done
cp ab/libmpc_new.so $SO; touch $SO
