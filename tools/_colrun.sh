MOG_BWD_IMPL=col python tools/dbg_bwd.py > gpurun_out/col_dbg.log 2>&1
MOG_BWD_IMPL=col python -m pytest tests/test_stn_gpu.py tests/test_stn_property_gpu.py -m gpu -x -q > gpurun_out/col_tests.log 2>&1; tail -3 gpurun_out/col_tests.log
for impl in auto col; do echo "== $impl"; MOG_BWD_IMPL=$impl python tools/kbench.py --cells 256:64:prior,256:28:prior,256:64:full,256:28:full,128:64:prior,128:28:prior,128:64:full,128:28:full,50:28:prior,50:28:full,64:28:prior,64:28:full --kinds write_bwd --tag col_$impl 2>&1 | tail -12; done > gpurun_out/col_kbench.log 2>&1
