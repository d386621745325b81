// ok_step_segstaged.cu -- the SEGMENT-STAGED shape of the beam kernel (ok_kernels.cuh: two 512-thread CTAs per SM, each
// behind its own track's header + segments in shared memory; centre line and grid read from the global blob) in a
// translation unit of its own, for the same reason as ok_step_unstaged.cu: kernels with different launch bounds must not
// share their out-of-line device functions.
#define OK_STEP_KERNEL_ONLY 1
#include "ok_kernels.cuh"

namespace ok
{
cudaError_t launch_step_segstaged(const StepParams &p, int grid, size_t smem_bytes, cudaStream_t stream)
{
    step_kernel<kBeamBlockSeg, true, true, true><<<grid, kBeamBlockSeg, smem_bytes, stream>>>(p, NoActor{});
    return cudaGetLastError();
}

// cudaFuncAttributeMaxDynamicSharedMemorySize = everything the device allows (an attribute of the function, not the env)
cudaError_t arm_step_segstaged(int smem_optin)
{
    cudaFuncAttributes fa{};
    cudaError_t        rc = cudaFuncGetAttributes(&fa, step_kernel<kBeamBlockSeg, true, true, true>);
    if (rc != cudaSuccess)
        return rc;
    return cudaFuncSetAttribute(step_kernel<kBeamBlockSeg, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                smem_optin - static_cast<int>(fa.sharedSizeBytes));
}

cudaError_t occupancy_step_segstaged(size_t smem_bytes, int *ctas_per_sm)
{
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, step_kernel<kBeamBlockSeg, true, true, true>, kBeamBlockSeg,
                                                         smem_bytes);
}

cudaError_t violations_step_segstaged(unsigned long long *count)
{
    return cudaMemcpyFromSymbol(count, g_violations, sizeof *count);
}
} // namespace ok
