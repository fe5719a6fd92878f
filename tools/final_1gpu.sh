#!/bin/bash
# Round-end measurement set on one B200 (run through gpurun): tests, smoke, bench (both arms), ncu launch list,
# full capture of one cold translation solve (copy-ring kernel + resumed tail).
T=${1:-r01d}
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; tail -1 gpurun_out/${T}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 400 python bench.py > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err; tail -c 300 gpurun_out/${T}_bench_1gpu.json; echo
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; tail -c 300 gpurun_out/${T}_bench_reference.json; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${T}_ncu_bench.log 2>&1; tail -c 200 gpurun_out/${T}_ncu_bench.log; echo
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_tsolve -s 8 -c 2 -f -o gpurun_out/${T}_tsolve python bench.py --steps 1 --warmup 0 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${T}_tsolve.log 2>&1; tail -2 gpurun_out/${T}_tsolve.log
