# round 2, sweep call (1 GPU): BASELINE configs 1, 3 and every C5 row on the final library (128 chunks per row)
mkdir -p gpurun_out
N_CHUNKS=128 timeout 560 python profiles/tools/config_sweep.py gpurun_out/r02_zz_config_sweep.md > gpurun_out/r02_zz_sweep.log 2>&1; tail -3 gpurun_out/r02_zz_sweep.log; wc -l gpurun_out/r02_zz_config_sweep.md
