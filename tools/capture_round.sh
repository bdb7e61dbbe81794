#!/bin/bash
# Evidence run of one round on a GPU box (through gpurun): GPU tests, bench line, launch lists, one ncu --set full capture
# per dominant kernel.  Every ncu command runs only after the same program exited 0 without ncu.  Outputs: gpurun_out/<tag>_*.
#   gpurun --timeout 1500 -- 'bash tools/capture_round.sh r02'
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
set -x
python __graft_entry__.py smoke > $out/${tag}_smoke.txt 2>&1 || exit 1
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_gputests.txt 2>&1
tail -3 $out/${tag}_gputests.txt
python bench.py --steps 10 --warmup 3 > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err || exit 2
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference_arm.json 2>/dev/null
python tools/probe_phases.py 6 147 10 12 14 16 17 18 19 20 22 > $out/${tag}_msm_phases.jsonl 2>&1
# launch lists (cold, serialised: shares only)
python tools/probe_phases.py --once 20 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_msm20.csv \
    python tools/probe_phases.py --once 20 > /dev/null 2>&1
python tools/probe_verify.py 1 > $out/${tag}_verify_plain.txt 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"verify_|vg_" -c 100 --csv --log-file $out/${tag}_launches_verify.csv \
    python tools/probe_verify.py 1 > /dev/null 2>&1
# full captures
ncu --set full --clock-control none --import-source on -k regex:msm_accumulate -c 8 -f -o $out/${tag}_acc \
    python tools/probe_phases.py --once 20 > $out/${tag}_ncu_acc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"vg_winsum|vg_fixed|vg_finish|vg_coeff" -c 8 -f -o $out/${tag}_verify \
    python tools/probe_verify.py 1 > $out/${tag}_ncu_verify.log 2>&1
ls -la $out | tail -20
