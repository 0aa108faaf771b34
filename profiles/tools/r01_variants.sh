# experiment driver: decompress kernel timings of the library variants named on the command line
mkdir -p gpurun_out
for v in "$@"; do if [ "$v" = default ]; then f=libcpcodec.so; else f=libcpcodec_$v.so; fi; PCOB200_LIB=$PWD/pcodec_b200/$f timeout 300 python profiles/tools/decompress_time.py 2>&1 | tail -1; done
