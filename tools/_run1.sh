set -x
for L in 1 2 3; do python bench.py --steps 5 --warmup 3 --no-cpu-baseline --lanes $L > gpurun_out/r2b_lanes$L.log 2> gpurun_out/r2b_lanes$L.err; done
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graphs --lanes 1 --profile gpurun_out/r2b_profile.md > gpurun_out/r2b_prof.log 2>&1
