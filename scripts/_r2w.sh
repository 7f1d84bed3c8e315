python -m pytest tests -m gpu -x -q -k "stations or ensemble or edge or dropin" > gpurun_out/r2w_tests.log 2>&1; echo tests rc=$?; tail -4 gpurun_out/r2w_tests.log
