#!/bin/bash
# Evidence pass of a round on ONE B200 (run through gpurun): bench lines for every BASELINE
# config, the reference / torch-eager arms, the ncu launch list of the bench command and one
# `ncu --set full` capture of a serial step at batch 8 and batch 64.  Output: gpurun_out/$1/
out=gpurun_out/${1:-evidence}
mkdir -p $out
for c in 5 2 1 3 4; do
  extra="--no-cpu-baseline"; [ $c = 5 ] && extra=""
  python bench.py --config $c $extra > $out/bench_c$c.json 2> $out/bench_c$c.err
done
python bench.py --config 5 --point-order sensor --no-cpu-baseline > $out/bench_c5_sensor_order.json 2> $out/bench_c5_sensor_order.err
python bench.py --impl reference --config 5 --steps 3 --warmup 1 > $out/bench_reference_c5.json 2> $out/ref.err
python bench.py --impl torch_eager --config 2 --steps 5 --warmup 2 > $out/bench_torch_eager_c2.json 2> $out/eager.err
for B in 8 64; do
  ncu --set full --clock-control none --import-source on --profile-from-start off -f -o $out/step_b$B \
      python tools/ncu_step.py $B > $out/ncu_b$B.log 2>&1
  ncu -i $out/step_b$B.ncu-rep --page raw --csv > $out/step_b${B}_raw.csv 2>> $out/ncu_b$B.log
  python tools/summarize_ncu.py $out/step_b${B}_raw.csv > $out/ncu_full_step_b$B.csv 2>> $out/ncu_b$B.log
  rm -f $out/step_b$B.ncu-rep      # 55 MB each: gpurun brings back at most 64 MiB
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_c5.csv \
    python bench.py --config 5 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --min-seconds 0 > $out/bench_under_ncu.log 2>&1
ls -la $out
