// Input side of the path: uint8 images -> the float32 (1, H, W) tensors the encoder takes.
// Mirrors the deterministic part of the reference's img_transform (data_wrangling/dataset.py:365-371):
//   ToTensor (x / 255) -> Grayscale(1) (0.2989 R + 0.587 G + 0.114 B, torchvision rgb_to_grayscale) -> Invert (1 - x),
// for a ragged batch in one launch, optionally zero-padding (= white background after Invert) every image to a multiple
// of the encoder's 16-pixel patch grid.  HBM-bound: 1 or 3 bytes read and 4 bytes written per pixel.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace {

// every operation rounded separately, in torchvision's order (no FMA contraction): bit-exact with the CPU pipeline
TX_DEVINL float inv_gray1(unsigned v) { return __fsub_rn(1.0f, __fdiv_rn((float)v, 255.f)); }
TX_DEVINL float inv_gray3(unsigned r8, unsigned g8, unsigned b8) {
    const float r = __fdiv_rn((float)r8, 255.f), g = __fdiv_rn((float)g8, 255.f), bl = __fdiv_rn((float)b8, 255.f);
    return __fsub_rn(1.0f, __fadd_rn(__fadd_rn(__fmul_rn(0.2989f, r), __fmul_rn(0.587f, g)), __fmul_rn(0.114f, bl)));
}

__global__ void __launch_bounds__(256) preprocess_u8_kernel(const uint8_t* __restrict__ in, const long* __restrict__ in_off,
                                                            const int* __restrict__ hwc, const long* __restrict__ out_off,
                                                            const int* __restrict__ out_hw, float* __restrict__ out) {
    const int b = blockIdx.y;
    const int H = hwc[3 * b], W = hwc[3 * b + 1], C = hwc[3 * b + 2];
    const int Hp = out_hw[2 * b], Wp = out_hw[2 * b + 1];
    const uint8_t* src = in + in_off[b];
    float* dst = out + out_off[b];
    const long n = (long)Hp * Wp;
    const long stride = (long)gridDim.x * blockDim.x;
    if ((W & 3) == 0 && (Wp & 3) == 0 && (((uintptr_t)src) & 3) == 0 && (((uintptr_t)dst) & 15) == 0) {
        // 4 pixels per thread: 4 or 12 source bytes as 32-bit words (rows start 4-byte aligned), one 16-byte store
        for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q < n / 4; q += stride) {
            const long i = q * 4;
            const int y = (int)(i / Wp), x = (int)(i - (long)y * Wp);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (y < H && x < W) {
                const uint32_t* w = reinterpret_cast<const uint32_t*>(src + ((long)y * W + x) * C);
                if (C == 1) {
                    const uint32_t a = __ldg(w);
                    v = make_float4(inv_gray1(a & 255), inv_gray1((a >> 8) & 255), inv_gray1((a >> 16) & 255), inv_gray1(a >> 24));
                } else {
                    const uint32_t a = __ldg(w), c = __ldg(w + 1), d = __ldg(w + 2);      // R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3
                    v.x = inv_gray3(a & 255, (a >> 8) & 255, (a >> 16) & 255);
                    v.y = inv_gray3(a >> 24, c & 255, (c >> 8) & 255);
                    v.z = inv_gray3((c >> 16) & 255, c >> 24, d & 255);
                    v.w = inv_gray3((d >> 8) & 255, (d >> 16) & 255, d >> 24);
                }
            }
            *reinterpret_cast<float4*>(dst + i) = v;
        }
        return;
    }
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int y = (int)(i / Wp), x = (int)(i - (long)y * Wp);
        float v = 0.f;
        if (y < H && x < W) {
            const uint8_t* px = src + ((long)y * W + x) * C;
            v = C == 1 ? inv_gray1(px[0]) : inv_gray3(px[0], px[1], px[2]);
        }
        dst[i] = v;
    }
}

}  // namespace

cudaError_t launch_preprocess_u8(const uint8_t* in, const long* in_off, const int* hwc, const long* out_off, const int* out_hw,
                                 float* out, int nimg, long max_out_pixels, cudaStream_t st) {
    if (nimg <= 0) return cudaSuccess;
    const unsigned bx = (unsigned)std::min<long>(std::max<long>((max_out_pixels / 4 + 255) / 256, 1), 1024);
    preprocess_u8_kernel<<<dim3(bx, nimg), 256, 0, st>>>(in, in_off, hwc, out_off, out_hw, out);
    return cudaGetLastError();
}
