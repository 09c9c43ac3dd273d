python tools/profile_sweep.py --width 1920 --height 1080 --passes 2 > /dev/null 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:rmd_eval_kernel -s 33 -c 33 \
  --csv --log-file gpurun_out/eval_traffic_r2b.csv python tools/profile_sweep.py --width 1920 --height 1080 --passes 2 > /dev/null 2>&1 && python tools/make_traffic.py gpurun_out/eval_traffic_r2b.csv r2b
timeout 300 python -m pytest tests -m gpu -x -q -k "golden_fixture or random_visits or extreme or full_1080p_sweep or prediction_samples or brief or cu_eval or 2160p" 2>&1 | tail -2
