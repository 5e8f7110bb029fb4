# compile only the headline instantiation of the step kernel and print its hot-loop instruction mix
set -e
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -DFDS_DEV_ONE -cubin -o /tmp/step_dev.cubin fds_b200/csrc/l96_step.cu
FDS_SASS_BIN=/tmp/step_dev.cubin python tools/sass_hot.py l96_rk4_stream_kernelINS_9ExactMath 900 "$@"
FDS_SASS_BIN=/tmp/step_dev.cubin python tools/sass_hot.py l96_rk4_stream_kernelINS_8FastMath 600 | head -3
