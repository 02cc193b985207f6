#!/bin/bash
# Everything under profiles/ that needs a GPU, in one gpurun call (tag = $1, default r02):
#   traffic of all 12 sweep cells (ncu, DRAM bytes per launch), the sweep itself, the headline cell's ncu --set full
#   capture and launch list.  Each ncu run follows a plain run of the same command that exited 0.
TAG=${1:-r02}
mkdir -p gpurun_out
NCU_M="ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:stn_(fwd|bwd) --csv"
rm -f gpurun_out/traffic_${TAG}_sweep.json
for cell in "50 28" "64 28" "128 28" "128 64" "256 28" "256 64"; do set -- $cell
  for regime in prior full; do
    CMD="python tools/traffic_capture.py run --canvas $1 --glimpse $2 --regime $regime"
    $CMD > gpurun_out/traffic_plain.log 2>&1 && $NCU_M --log-file gpurun_out/traffic_$1_$2_$regime.csv $CMD > gpurun_out/traffic_ncu.log 2>&1 \
      && python tools/traffic_capture.py parse gpurun_out/traffic_$1_$2_$regime.csv --canvas $1 --glimpse $2 --regime $regime --into gpurun_out/traffic_${TAG}_sweep.json \
      && echo "traffic $1 $2 $regime ok" || { echo "traffic $1 $2 $regime FAILED"; tail -3 gpurun_out/traffic_ncu.log; }
  done
done
mkdir -p profiles && cp gpurun_out/traffic_${TAG}_sweep.json profiles/traffic_${TAG}_sweep.json
python tools/traffic_capture.py parse gpurun_out/traffic_256_64_prior.csv --canvas 256 --glimpse 64 --regime prior > gpurun_out/traffic_latest.json \
  && cp gpurun_out/traffic_latest.json profiles/traffic_latest.json
python bench.py --sweep --steps 3 --tag ${TAG} > gpurun_out/sweep_${TAG}.jsonl 2> gpurun_out/sweep_${TAG}.err; echo "sweep rc=$?"
cp profiles/sweep_${TAG}.json gpurun_out/sweep_${TAG}.json
# headline cell: full capture of the four kernels (first launch of each kind) and the launch list of a short run
CMD="python tools/traffic_capture.py run --canvas 256 --glimpse 64 --regime prior"
$CMD > gpurun_out/full_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k 'regex:stn_(fwd|bwd)' -c 4 -o gpurun_out/prof_${TAG}_256x64 $CMD > gpurun_out/full_ncu.log 2>&1; echo "full rc=$?"
CMD="python tools/traffic_capture.py run --canvas 50 --glimpse 28 --regime prior"
$CMD > gpurun_out/full_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k 'regex:stn_(fwd|bwd)' -c 4 -o gpurun_out/prof_${TAG}_50x28 $CMD > gpurun_out/full_ncu2.log 2>&1; echo "full2 rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-train"
$CMD > gpurun_out/ll_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}_256x64.csv $CMD > gpurun_out/ll_ncu.log 2>&1; echo "launch list rc=$?"
python tools/launch_list.py gpurun_out/launches_${TAG}_256x64.csv "ncu --metrics gpu__time_duration.sum --clock-control none -c 400 : $CMD (headline cell: canvas 256, glimpse 64, prior-like theta, B = 16384)" > gpurun_out/${TAG}_stn_c5_256x64_launch_list.txt
for c in 256x64 50x28; do python tools/ncu_summary.py gpurun_out/prof_${TAG}_$c.ncu-rep > gpurun_out/${TAG}_stn_c5_${c}_ncu_full.txt 2>/dev/null; done
