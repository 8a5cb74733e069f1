set -x
T=$1
python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1
timeout 300 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 400 python bench.py --impl reference > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
for c in 1 3 4; do timeout 300 python bench.py --config $c > gpurun_out/${T}_cfg$c.json 2> gpurun_out/${T}_cfg$c.err; done
timeout 300 python bench.py --config 4 --cfg4-mode file --no-cpu-baseline > gpurun_out/${T}_cfg4file.json 2> gpurun_out/${T}_cfg4file.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_ncu1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:stft_frames|yin_fft|conv1_pool|conv_tc" -s 14 -c 7 -o gpurun_out/${T}_full -f python tools/stft_only.py full > gpurun_out/${T}_ncu2.log 2>&1
timeout 400 python tools/parity_report.py > gpurun_out/${T}_parity.json 2> gpurun_out/${T}_parity.err
ls -la gpurun_out/${T}_*
