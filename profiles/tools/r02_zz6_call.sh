# round 2: ncu --set full of walk_kernel (1024 chunks walked at once, one thread each) for DESIGN.md 9.4
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'walk_kernel' -s 1 -c 1 -f -o gpurun_out/r02_zz6_walk python profiles/tools/walk_once.py > gpurun_out/r02_zz6_ncu.log 2>&1; tail -2 gpurun_out/r02_zz6_ncu.log
