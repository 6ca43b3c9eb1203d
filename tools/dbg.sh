python -m pytest tests/test_gpu_kernels_basic.py -q -k staged 2>&1 | grep -E "Error|error" | head -5
timeout 200 python -X faulthandler -c "
import faulthandler, sys, runpy
faulthandler.dump_traceback_later(100, exit=True)
sys.argv=['bench.py','--no-cpu-baseline','--no-pipe-peaks','--no-configs','--steps','5','--warmup','3']
runpy.run_path('bench.py', run_name='__main__')
" > gpurun_out/r02e_dbg.json 2> gpurun_out/r02e_dbg.err
tail -40 gpurun_out/r02e_dbg.err
