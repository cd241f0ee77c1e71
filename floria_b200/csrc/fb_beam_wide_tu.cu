// fb_beam_wide_tu.cu — translation unit of k_beam_wide (one large beam-search instance spread over the whole GPU).
// Compiled in parallel with fb_lib.cu and fb_beam_tu.cu (floria_b200/build.py).
#define FB_BEAM_WIDE_IMPL
#include "fb_beam_wide.cuh"

// largest cooperative grid (one CTA per SM at most: the slices are dealt round-robin, more CTAs per SM would only
// lengthen the grid barrier)
int fb_beam_wide_max_grid(size_t smem_bytes, int sm_count, int *grid) {
    cudaError_t e = cudaFuncSetAttribute(k_beam_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return (int)e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_beam_wide, FB_BW_THREADS, smem_bytes);
    if (e != cudaSuccess) return (int)e;
    *grid = per_sm >= 1 ? sm_count : 0;
    return 0;
}

int fb_beam_wide_launch(unsigned grid, size_t smem_bytes, cudaStream_t stream, const BeamParams &bp) {
    void *args[] = {(void *)&bp};
    cudaError_t e = cudaLaunchCooperativeKernel((const void *)k_beam_wide, dim3(grid), dim3(FB_BW_THREADS), args, smem_bytes, stream);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaGetLastError();
}
