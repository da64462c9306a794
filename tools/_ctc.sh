python -m pytest tests/test_ctc_gpu.py -m gpu -x -q > gpurun_out/r2h_ctc_tests.log 2>&1; tail -12 gpurun_out/r2h_ctc_tests.log
python - <<'P' > gpurun_out/r2h_ctc_sweep.log 2>&1
import json, torch, sys
sys.path.insert(0, '.')
import bench
from metaasr_crossaccent_b200 import ops
be = ops.CudaBackend(torch.device('cuda', 0), torch.bfloat16, gemm='umma')
r = bench.ctc_bandwidth(be, bench.load_peaks(), full=True)
for k, v in r.items(): print(k, v)
P
cat gpurun_out/r2h_ctc_sweep.log
