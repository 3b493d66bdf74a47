mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu -x --timeout 120 2>&1 | tail -2
timeout 200 python bench.py --steps 40 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_final.json
cut -c1-200 gpurun_out/bench_final.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r01_launches_bench.log 2>&1
for l in conv1 deconv3; do timeout 200 ncu --set full --import-source on --clock-control none -k regex:k_conv_tc -s 3 -c 1 -f -o gpurun_out/r01_${l}_full python tools/prof_layers.py $l 2 > gpurun_out/r01_${l}_full.log 2>&1; done
timeout 100 python tools/bench_configs.py all 2>&1 | tail -5 > gpurun_out/bench_configs.jsonl
cut -c1-120 gpurun_out/bench_configs.jsonl
