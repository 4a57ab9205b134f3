mkdir -p gpurun_out/r02o
( timeout 900 python -m pytest tests/test_gpu_graph_build.py tests/test_gpu_mds.py tests/test_gpu_dropin.py -q ) > gpurun_out/r02o/pytest.log 2>&1; tail -12 gpurun_out/r02o/pytest.log
