// tile_inst.cu — one instantiation of tile_kernel<FS, S> for every sample count, compiled once per fragment shader id
// (-DIFR_INST_FS=<id>, see the Makefile).
#include "tile_kernel.cuh"
#define IFR_CAT2(a, b) a##b
#define IFR_CAT(a, b) IFR_CAT2(a, b)
cudaError_t IFR_CAT(ifr_launch_fs_, IFR_INST_FS)(const DrawParams& p, cudaStream_t s) { return launch_fs<IFR_INST_FS>(p, s); }
