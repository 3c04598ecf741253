#!/bin/bash
# One gpurun call that produces the evidence profiles/ keeps: GPU parity tests, the bench lines of every config, the launch list of
# the headline bench and one `ncu --set full` capture per hot kernel.  Usage (from the repo root, on the GPU box):
#   tools/gpu_round.sh [tests] [bench] [launches] [ncu:bc7] [ncu:bc7tma] [ncu:bc6h] [ncu:cubic] [ncu:linear] [ncu:bc3] [ncu:convert] [prof]
# Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
want() { for a in "$@"; do for b in $ARGS; do [ "$a" = "$b" ] && return 0; done; done; return 1; }
ARGS="${*:-tests bench launches ncu:bc7 ncu:bc7tma ncu:bc6h ncu:cubic ncu:linear ncu:bc3 ncu:convert prof}"

if want tests; then
    timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
    tail -3 $O/pytest_gpu.log
fi
if want prof; then
    timeout 600 python tools/prof_driver.py all 10 > $O/prof_driver_timings.txt 2>&1
    for m in 1 2 3; do echo "DXB200_BC7_TMA=$m" >> $O/prof_driver_timings.txt; DXB200_BC7_TMA=$m timeout 200 python tools/prof_driver.py bc7 10 >> $O/prof_driver_timings.txt 2>&1; done
    cat $O/prof_driver_timings.txt
fi
if want bench; then
    : > $O/bench_lines.jsonl
    for c in c2 c3 c4 c5; do
        timeout 900 python bench.py --config $c --steps 20 --warmup 3 2> $O/bench_$c.err | tail -1 >> $O/bench_lines.jsonl
    done
    timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>> $O/bench_c2.err | tail -1 >> $O/bench_lines.jsonl
    cut -c1-400 $O/bench_lines.jsonl
fi
if want launches; then
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_bc7.csv \
        python bench.py --steps 2 --warmup 3 --batch 2 > $O/launches_bench.log 2>&1
fi
cap() {  # name kernel-regex skip driver-args TU units note   (the summary is written on the box; reports above 12 MB are dropped: gpurun_out/ is capped at 64 MiB)
    timeout 900 $NCU -k "regex:$2" -c 1 -s $3 -f -o $O/$1 python tools/prof_driver.py $4 1 > $O/$1.log 2>&1
    python tools/ncu_summary.py $O/$1.ncu-rep $5 $6 "ncu --set full --clock-control none --import-source on -k regex:$2 -c 1 -s $3 python tools/prof_driver.py $4 1 | $7" > $O/$1.txt 2>> $O/$1.log
    [ $(stat -c %s $O/$1.ncu-rep) -gt 12000000 ] && rm -f $O/$1.ncu-rep
}
want ncu:bc7 && cap r02_ncu_k_compress_bc7 k_compress_bc7 2 bc7 dxb_k_bc7 1048576 "4096x4096 RGBA32F -> BC7_UNORM (1,048,576 blocks), B200, round 2"
want ncu:bc7tma && DXB200_BC7_TMA=1 cap r02_ncu_k_compress_bc7_tma k_compress_bc7_tma 2 bc7 dxb_k_bc7 1048576 "DXB200_BC7_TMA=1: 4096x4096 RGBA32F -> BC7_UNORM through the TMA-fed persistent kernel (opt-in), B200, round 2"
want ncu:bc6h && cap r02_ncu_k_compress_bc6h k_compress_bc6h 2 bc6h dxb_k_bc6h 262144 "2048x2048 RGBA16F -> BC6H_UF16 (262,144 blocks), B200, round 2"
want ncu:cubic && cap r02_ncu_k_mip_sep_cubic k_mip_sep 0 rowscubic dxb_k_rows 16777216 "64 x 1024^2 RGBA8 CUBIC chain, first level (64 x 512^2 outputs), B200, round 2"
want ncu:linear && cap r02_ncu_k_mip_box3_linear k_mip_box3 0 rowslinear dxb_k_rows 16777216 "64 x 1024^2 RGBA8 LINEAR chain, first launch = levels 1-3 fused (k_mip_box3<..., LIN>), B200, round 2"
want ncu:bc3 && cap r02_ncu_c4 k_compress_bc15 2 bc3 dxb_k_bc15 1048576 "4096x4096 RGBA8 -> BC3_UNORM (1,048,576 blocks; the kernel of C4), B200, round 2"
want ncu:convert && cap r02_ncu_k_convert_vec k_convert_vec 2 rows dxb_k_rows 67108864 "8192x8192 R8 -> R32F (C5), B200, round 2"
ls -la $O
