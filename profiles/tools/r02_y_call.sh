# round 2, call y (1 GPU): phase trace (PCOB200_TRACE) of the streamed e2e leg with the FIFO copy policy, G=16 P=2 and G=32 P=3 (16 MB slices)
mkdir -p gpurun_out
PCOB200_TRACE=1 timeout 300 python profiles/tools/e2e_in_place_probe.py 1024 0 16x2 > gpurun_out/r02_y_trace_16x2.out 2> gpurun_out/r02_y_trace_16x2.err
PCOB200_TRACE=1 PCOB200_COPY_SLICE_MB=16 timeout 300 python profiles/tools/e2e_in_place_probe.py 1024 0 32x3 > gpurun_out/r02_y_trace_32x3.out 2> gpurun_out/r02_y_trace_32x3.err
for v in 16x2 32x3; do
  awk '/=== timed passes/{f=1} f' gpurun_out/r02_y_trace_$v.err | head -80 > gpurun_out/r02_y_trace_$v.txt
  cat gpurun_out/r02_y_trace_$v.out | tail -2
done
