// fb_beam_tu.cu — translation unit of the beam-search kernel (k_beam<256>, k_beam<128>, seven ploidies each).
// Kept apart from fb_lib.cu so that the two compile in parallel (floria_b200/build.py).
#include "fb_beam.cuh"

template <int NT>
static int fb_beam_prepare(size_t smem_bytes) {
    return (int)cudaFuncSetAttribute(k_beam<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
}

int fb_beam_occupancy(int threads, size_t smem_bytes, int *blocks_per_sm) {
    cudaError_t e;
    if (threads == FB_BEAM_THREADS_SMALL) {
        if ((e = (cudaError_t)fb_beam_prepare<FB_BEAM_THREADS_SMALL>(smem_bytes)) != cudaSuccess) return (int)e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k_beam<FB_BEAM_THREADS_SMALL>, threads, smem_bytes);
    } else {
        if ((e = (cudaError_t)fb_beam_prepare<FB_BEAM_THREADS>(smem_bytes)) != cudaSuccess) return (int)e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k_beam<FB_BEAM_THREADS>, threads, smem_bytes);
    }
    return (int)e;
}

int fb_beam_launch(int threads, unsigned grid, size_t smem_bytes, cudaStream_t stream, const BeamParams &bp) {
    if (threads == FB_BEAM_THREADS_SMALL)
        k_beam<FB_BEAM_THREADS_SMALL><<<grid, threads, smem_bytes, stream>>>(bp);
    else
        k_beam<FB_BEAM_THREADS><<<grid, threads, smem_bytes, stream>>>(bp);
    return (int)cudaGetLastError();
}
