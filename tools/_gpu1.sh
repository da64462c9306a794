set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2i_tests.log 2>&1; tail -4 gpurun_out/r2i_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2i_bench.log 2> gpurun_out/r2i_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2i_ref.log 2> gpurun_out/r2i_ref.err
# sanitizer: hand-rolled mbarrier / flag protocols
S="compute-sanitizer --print-limit 20"
$S --tool memcheck python -m pytest tests/test_ctc_gpu.py -m gpu -x -q -k "128-32-367-32 or edge" > gpurun_out/r2i_san_ctc_memcheck.log 2>&1; tail -3 gpurun_out/r2i_san_ctc_memcheck.log
$S --tool racecheck python -m pytest tests/test_ctc_gpu.py -m gpu -x -q -k "128-32-367-32" > gpurun_out/r2i_san_ctc_racecheck.log 2>&1; tail -3 gpurun_out/r2i_san_ctc_racecheck.log
$S --tool racecheck python -m pytest tests/test_conv_umma_gpu.py -m gpu -x -q -k "2-64-83-64-64" > gpurun_out/r2i_san_conv_racecheck.log 2>&1; tail -3 gpurun_out/r2i_san_conv_racecheck.log
$S --tool racecheck python -m pytest tests/test_attn_umma_gpu.py -m gpu -x -q -k "3-8-33-128 or 2-4-70-300" > gpurun_out/r2i_san_attn_racecheck.log 2>&1; tail -3 gpurun_out/r2i_san_attn_racecheck.log
$S --tool memcheck python -m pytest tests/test_attn_umma_gpu.py tests/test_conv_umma_gpu.py -m gpu -x -q > gpurun_out/r2i_san_umma_memcheck.log 2>&1; tail -3 gpurun_out/r2i_san_umma_memcheck.log
