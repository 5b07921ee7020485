// Input side of the path: uint8 images -> the float32 (1, H, W) tensors the encoder takes.
// Mirrors the deterministic part of the reference's img_transform (data_wrangling/dataset.py:365-371):
//   ToTensor (x / 255) -> Grayscale(1) (0.2989 R + 0.587 G + 0.114 B, torchvision rgb_to_grayscale) -> Invert (1 - x),
// for a ragged batch in one launch, optionally zero-padding (= white background after Invert) every image to a multiple
// of the encoder's 16-pixel patch grid.  HBM-bound: 1 or 3 bytes read and 4 bytes written per pixel.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace {

__global__ void __launch_bounds__(256) preprocess_u8_kernel(const uint8_t* __restrict__ in, const long* __restrict__ in_off,
                                                            const int* __restrict__ hwc, const long* __restrict__ out_off,
                                                            const int* __restrict__ out_hw, float* __restrict__ out) {
    const int b = blockIdx.y;
    const int H = hwc[3 * b], W = hwc[3 * b + 1], C = hwc[3 * b + 2];
    const int Hp = out_hw[2 * b], Wp = out_hw[2 * b + 1];
    const uint8_t* src = in + in_off[b];
    float* dst = out + out_off[b];
    const long n = (long)Hp * Wp;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int y = (int)(i / Wp), x = (int)(i - (long)y * Wp);
        float v = 0.f;
        if (y < H && x < W) {
            const uint8_t* px = src + ((long)y * W + x) * C;
            float gray;
            // every operation rounded separately, in torchvision's order (no FMA contraction): bit-exact with the CPU pipeline
            if (C == 1) gray = __fdiv_rn((float)px[0], 255.f);
            else {
                const float r = __fdiv_rn((float)px[0], 255.f), g = __fdiv_rn((float)px[1], 255.f), bl = __fdiv_rn((float)px[2], 255.f);
                gray = __fadd_rn(__fadd_rn(__fmul_rn(0.2989f, r), __fmul_rn(0.587f, g)), __fmul_rn(0.114f, bl));
            }
            v = __fsub_rn(1.0f, gray);
        }
        dst[i] = v;
    }
}

}  // namespace

cudaError_t launch_preprocess_u8(const uint8_t* in, const long* in_off, const int* hwc, const long* out_off, const int* out_hw,
                                 float* out, int nimg, long max_out_pixels, cudaStream_t st) {
    if (nimg <= 0) return cudaSuccess;
    const unsigned bx = (unsigned)std::min<long>(std::max<long>((max_out_pixels + 255) / 256, 1), 1024);
    preprocess_u8_kernel<<<dim3(bx, nimg), 256, 0, st>>>(in, in_off, hwc, out_off, out_hw, out);
    return cudaGetLastError();
}
