#!/bin/bash
# turns an .ncu-rep into text that fits gpurun_out (details page, raw CSV, gzip'd source CSV) and removes the report
rep=$1; base=${rep%.ncu-rep}
ncu -i $rep --page details > ${base}_details.txt 2>&1
ncu -i $rep --page raw --csv > ${base}_raw.csv 2>&1
ncu -i $rep --page source --csv 2>/dev/null | gzip -9 > ${base}_source.csv.gz
rm -f $rep
ls -la ${base}_*
