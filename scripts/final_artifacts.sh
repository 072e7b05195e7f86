V=v42
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests_$V.txt 2>&1; tail -1 gpurun_out/r02_gpu_tests_$V.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_$V.txt 2>&1; tail -1 gpurun_out/r02_smoke_$V.txt
python bench.py > gpurun_out/r02_bench_n1_$V.json 2> gpurun_out/bench_$V.err; echo bench rc=$?
python bench.py --impl reference > gpurun_out/r02_bench_reference_arm_$V.json 2>> gpurun_out/bench_$V.err; echo ref rc=$?
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
$NCU -c 700 --log-file gpurun_out/r02_launches_$V.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --parity "" --parity-modes fp32tc --testset-views 0 --train-steps 0 --no-config5 > /dev/null 2>&1; echo l1 rc=$?
$NCU -c 600 --log-file gpurun_out/r02_launches_train_$V.csv python scripts/train_once.py 14 > /dev/null 2>&1; echo l2 rc=$?
$NCU -c 400 --log-file gpurun_out/r02_launches_skip_$V.csv python scripts/skip_once.py 1 > /dev/null 2>&1; echo l3 rc=$?
$NCU -c 400 --log-file gpurun_out/r02_launches_kilo_$V.csv python scripts/kilo_once.py 0 32 > /dev/null 2>&1; echo l4 rc=$?
python scripts/train_once_compat.py > gpurun_out/r02_train_variants_$V.txt 2>&1; tail -3 gpurun_out/r02_train_variants_$V.txt
python scripts/kilo_ab.py gpurun_out/kilo_tc.npy
