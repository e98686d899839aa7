#!/bin/bash
# usage: ncu_export.sh <report.ncu-rep> : raw / source / details pages as text next to the report, then the report is removed
# (gpurun copies at most 64 MiB back; a --set full report of one launch is ~30 MB)
set -e
r="$1"; b="${r%.ncu-rep}"
ncu -i "$r" --page raw --csv > "${b}_raw.csv" 2>/dev/null
ncu -i "$r" --page source --csv > "${b}_src.csv" 2>/dev/null
ncu -i "$r" --page details > "${b}_details.txt" 2>/dev/null
rm -f "$r"
