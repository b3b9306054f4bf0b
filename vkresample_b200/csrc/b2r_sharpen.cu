// b2r_sharpen.cu -- K8 launcher (CAS-style sharpen, see b2r_kernels.cuh).
#include "b2r_launch.h"

namespace b2r {
cudaError_t launch_sharpen_kernel(cudaStream_t s, const SharpenArgs& a) {
    const int bx = sharpen_rows_block(a.dm.up_w);
    if (bx > 0) {   // vectorised rolling-window kernel
        constexpr int RY = kSharpenRowsPerThread;
        dim3 block(bx), grid(a.dm.up_w / 4 / bx, (a.dm.up_h + RY - 1) / RY, 3);
        if (a.precision == 2)
            k_sharpen_rows<__half, RY><<<grid, block, 0, s>>>((const __half*)a.pre, (__half*)a.out, a.dm);
        else
            k_sharpen_rows<float, RY><<<grid, block, 0, s>>>((const float*)a.pre, (float*)a.out, a.dm);
        return cudaGetLastError();
    }
    constexpr int PX = 4;   // any-width fallback
    dim3 block(256), grid((a.dm.up_w + PX * 256 - 1) / (PX * 256), a.dm.up_h, 3);
    if (a.precision == 2)
        k_sharpen<__half, PX><<<grid, block, 0, s>>>((const __half*)a.pre, (__half*)a.out, a.dm);
    else
        k_sharpen<float, PX><<<grid, block, 0, s>>>((const float*)a.pre, (float*)a.out, a.dm);
    return cudaGetLastError();
}
}  // namespace b2r
