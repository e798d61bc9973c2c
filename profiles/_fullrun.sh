mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "=== bench"; timeout 400 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_line.json | cut -c1-400
echo "=== bench reference"; timeout 400 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference_line.json | cut -c1-300
echo "=== ncu autodetect"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/autodetect.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
echo "=== launches"; timeout 600 bash profiles/run_launches.sh r01 2>&1 | tail -14
echo "=== ncu full"; timeout 900 bash profiles/run_ncu_full.sh r01 2>&1 | tail -3
python profiles/summarize_ncu.py r01 2>&1 | tail -3
for w in c1 c2 c4 c5; do echo "=== $w"; timeout 200 python profiles/quick_pair.py $w 300 2>&1 | tail -1 | cut -c1-260; done
echo "=== smoke"; timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
