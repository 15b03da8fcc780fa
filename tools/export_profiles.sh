#!/usr/bin/env bash
# Regenerate the committed profiling evidence under profiles/ from ncu reports and from the built library.
#
#   tools/export_profiles.sh <tag> [report.ncu-rep ...]
#
# For every report: profiles/<tag>_<report>_raw.csv        (ncu --page raw: every metric of every captured launch)
#                   profiles/<tag>_<report>_summary.txt     (the metrics the READMEs quote, one block per launch)
# From zutis_b200/libzutis_b200.so: profiles/sass_<kernel>.txt (cuobjdump -sass of the hot kernels, mnemonics only) and
#                   profiles/sass_mnemonics.txt             (counts of the Blackwell-specific mnemonics per kernel)
# The reports themselves are made on the GPU box, e.g. (see /opt/skills/guides/B200_PROFILING.md):
#   ncu --set full --clock-control none --import-source on -k regex:'decode_cells|gemm_tcgen05' -s 6 -c 2 \
#       -o gpurun_out/<name> python bench.py --steps 4 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-extras
#   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/<name>_launches.csv \
#       python bench.py --steps 40 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-extras
set -euo pipefail
cd "$(dirname "$0")/.."
tag=${1:?usage: tools/export_profiles.sh <tag> [report.ncu-rep ...]}
shift || true
mkdir -p profiles
for rep in "$@"; do
    name=$(basename "$rep" .ncu-rep)
    ncu -i "$rep" --page raw --csv > "profiles/${tag}_${name}_raw.csv"
    python - "$rep" "profiles/${tag}_${name}_raw.csv" > "profiles/${tag}_${name}_summary.txt" <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[2])))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_active.avg", "sm__cycles_active.max", "sm__cycles_elapsed.max"]
print("source report:", sys.argv[1])
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")], "| launch id", r[hdr.index("ID")])
    for w in want:
        if w in hdr:
            print(f"   {w:72s} {r[hdr.index(w)]:>16s} {units[hdr.index(w)]}")
PY
    echo "profiles/${tag}_${name}_raw.csv profiles/${tag}_${name}_summary.txt"
done
lib=zutis_b200/libzutis_b200.so
: > profiles/sass_mnemonics.txt
# kernel name -> substring of the mangled name that selects the instantiation (decode_cells: int64 ground truth, NI = 3,
# the one the headline workload launches; the others: the first instantiation in the library)
for spec in gemm_tcgen05_kernel:gemm_tcgen05_kernel decode_cells_kernel:decode_cells_kernelIxLi3 decode_tiled_kernel:decode_tiled_kernel \
            threshold_tiled_kernel:threshold_tiled_kernel nms_hard_kernel:nms_hard_kernel; do
    k=${spec%%:*}; pat=${spec##*:}
    out="profiles/sass_${k}.txt"
    # mnemonics and operands only (no encodings)
    cuobjdump -sass "$lib" | awk -v k="$pat" '
        /Function :/ { if (on) exit; if (index($0, k) > 0) { on = 1; print } next }
        on && /^[[:space:]]+\/\*[0-9a-f]{4}\*\// { sub(/\/\* 0x[0-9a-f]+ \*\//, ""); sub(/[[:space:]]+$/, ""); print }' > "$out"
    {
        echo "== $k ($(grep -c '/\*' "$out" || true) instructions; $(head -1 "$out" | sed -e 's/^[[:space:]]*//'))"
        for m in UTCHMMA UTCQMMA UTMALDG UTMASTG UBLKCP LDTM STTM UTCBAR SYNCS FFMA2 FMUL2 FADD2 FMNMX3 MATCH.ANY REDUX ATOMS RED LDGSTS; do
            c=$(grep -c "[[:space:]]$m" "$out" || true)
            [ "$c" != "0" ] && echo "   $m: $c"
        done
    } >> profiles/sass_mnemonics.txt
done
echo "profiles/sass_*.txt profiles/sass_mnemonics.txt"
