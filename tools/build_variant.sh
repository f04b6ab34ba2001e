# dev helper: build a variant of libake_b200.so with extra -D flags into gpurun_out-independent tools/bin/<name>.so
# usage: bash tools/build_variant.sh <name> -DAKE_P2P_GROUPS=4 ...
name=$1; shift
mkdir -p tools/bin
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O2 --shared -cudart shared "$@" -I include \
  -o tools/bin/$name.so audio_key_estimation_b200/csrc/pcn.cu audio_key_estimation_b200/csrc/cqt.cu audio_key_estimation_b200/csrc/pipeline.cu -Xptxas -v 2>&1 | grep -A2 "p2p_umma_kernelILb0ELb0" | grep -E "registers|spill"
