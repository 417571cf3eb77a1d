#!/bin/bash
# The measurement set profiles/r2_* is made from: one box, the shipped build, nothing under a profiler except the launch list.
O=gpurun_out/final; mkdir -p $O
python bench.py --impl reference --steps 20 --warmup 5 > $O/ref.json 2> $O/ref.err
python bench.py --steps 20 --warmup 5 > $O/n1.json 2> $O/n1.err
python bench.py --precision f32 --steps 20 --warmup 5 > $O/n1_f32.json 2> $O/n1_f32.err
: > $O/configs.jsonl
for p in f64 f32; do for c in 1 2 3 4; do python bench.py --precision $p --config $c --steps 20 --warmup 5 >> $O/configs.jsonl 2>> $O/configs.err; done; done
: > $O/kernels.jsonl
for n in 2048 4096 1024 16384; do h=$((n/4)); [ $n = 2048 ] && h=512; python tools/kbench.py --tag r2final --n $n --hop $h >> $O/kernels.jsonl 2>> $O/kernels.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 > $O/ncu.log 2>&1
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
head -c 600 $O/n1.json
