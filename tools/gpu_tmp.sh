mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-knn > gpurun_out/launches_r1f.log 2>&1
tail -3 gpurun_out/launches_r1f.log | cut -c1-300
