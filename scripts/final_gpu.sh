mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/gputest_final.log 2>&1; tail -3 gpurun_out/gputest_final.log
timeout 600 python bench.py > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/launches_r02b.csv python bench.py --config 2 --steps 2 --no-extra --no-cpu-baseline > gpurun_out/ncu_ll.log 2>&1
for k in rnn_fwd_kernel rnn_bwd_kernel dw_mm16_kernel bulk_dp_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -o gpurun_out/full_$k -f python bench.py --config 2 --steps 2 --no-extra --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fft_features_kernel -s 4 -c 1 -o gpurun_out/full_fft_features_kernel -f python scripts/run_fft.py > gpurun_out/ncu_fft.log 2>&1
timeout 200 python scripts/dbg_rnn_fwd.py 3 64 > gpurun_out/tl_fwd_final.txt 2>&1
timeout 200 python scripts/dbg_rnn_bwd.py 3 > gpurun_out/tl_bwd_final.txt 2>&1
ls -la gpurun_out/*.ncu-rep
