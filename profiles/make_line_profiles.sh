#!/bin/bash
# usage: profiles/make_line_profiles.sh <report.ncu-rep> <out-dir> <label>
# per-source-line totals of the batched kernels of one `ncu --set full --import-source on` report
REP=$(readlink -f $1); OUT=$(readlink -f $2); LABEL=$3
HERE=$(dirname $(readlink -f $0))
cd /tmp
ncu -i $REP --page raw --csv > $OUT/${LABEL}_raw.csv 2>/dev/null
for k in assoc_cells_batch extract_select_batch extract_normals_rows eval_global moment_batch segment_scatter_batch map_cells_batch map_insert_batch; do
  ncu -i $REP --page source --print-source cuda,sass --csv -k regex:$k > /tmp/_src_$k.csv 2>/dev/null
  python $HERE/line_profile.py /tmp/_src_$k.csv 28 | cut -c1-190 > $OUT/${LABEL}_${k}_lines.txt
done
