// What does one link of a dependent-kernel chain cost on B200?  A chain of NK small kernels (each CTA reads what the
// previous kernel wrote, adds one, writes it back) is captured into a CUDA graph and replayed; the period per kernel is
// reported for three ways of ordering consecutive kernels:
//   plain : ordinary stream order (full launch + drain per kernel)
//   pdl   : programmatic dependent launch, early trigger, griddepcontrol.wait before the first read
//   flag  : programmatic dependent launch for early residency, but the data dependency is a release/acquire counter in
//           global memory that the producer's CTAs bump after their stores (no wait for grid completion)
// Several independent chains can run side by side on their own streams (like the decode branches).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pdl_chain_probe pdl_chain_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// mode 0: plain, 1: pdl, 2: flag.  buf: [2][n] ping-pong; cnt: per-link counters; epoch: replay number (device word)
__global__ void link_kernel(int mode, const float* in, float* out, int n, const unsigned* wait_cnt, unsigned wait_ctas,
                            unsigned* my_cnt, const unsigned* epoch, int* err) {
    if (mode >= 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (mode == 1) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (mode == 2) {
        if (wait_cnt) {
            if (threadIdx.x == 0) {
                const unsigned target = wait_ctas * (*epoch + 1);      // epoch is only bumped between graph replays
                const unsigned long long t0 = gtime();
                while (ld_acquire(wait_cnt) < target) {
                    if (gtime() - t0 > 20000000ull) { *err = 1; break; }      // 20 ms: never hang the box
                }
            }
            __syncthreads();
        } else {
            asm volatile("griddepcontrol.wait;" ::: "memory");      // first link of a replay: wait for the previous replay
        }
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __ldcg(in + (i + 32) % n) + 1.0f;            // reads a neighbour written by another CTA of the predecessor
    if (mode == 2) {
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(my_cnt, 1u);
    }
}
__global__ void bump_epoch(unsigned* epoch) { *epoch += 1; }

int main(int argc, char** argv) {
    const int NK = 40, CTAS = argc > 1 ? atoi(argv[1]) : 16, THREADS = 128, REPS = 200;
    const int n = CTAS * THREADS;
    for (int chains : {1, 6}) {
        for (int mode = 0; mode < 3; ++mode) {
            std::vector<cudaStream_t> st(chains);
            std::vector<cudaGraphExec_t> ge(chains);
            std::vector<float*> buf(chains);
            std::vector<unsigned*> cnt(chains);
            int* err; CK(cudaMalloc(&err, 4)); CK(cudaMemset(err, 0, 4));
            for (int c = 0; c < chains; ++c) {
                CK(cudaStreamCreateWithFlags(&st[c], cudaStreamNonBlocking));
                CK(cudaMalloc(&buf[c], 2 * n * sizeof(float))); CK(cudaMemset(buf[c], 0, 2 * n * sizeof(float)));
                CK(cudaMalloc(&cnt[c], (NK + 1) * 4)); CK(cudaMemset(cnt[c], 0, (NK + 1) * 4));
                unsigned* epoch = cnt[c] + NK;
                CK(cudaDeviceSynchronize());
                cudaGraph_t g;
                CK(cudaStreamBeginCapture(st[c], cudaStreamCaptureModeRelaxed));
                for (int k = 0; k < NK; ++k) {
                    cudaLaunchConfig_t cfg{};
                    cfg.gridDim = dim3(CTAS); cfg.blockDim = dim3(THREADS); cfg.stream = st[c];
                    cudaLaunchAttribute at[1];
                    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                    at[0].val.programmaticStreamSerializationAllowed = 1;
                    cfg.attrs = at; cfg.numAttrs = (mode >= 1 && k > 0) ? 1 : 0;
                    const float* in = buf[c] + (k & 1) * n; float* out = buf[c] + ((k + 1) & 1) * n;
                    const unsigned* wc = k > 0 ? cnt[c] + (k - 1) : nullptr;
                    CK(cudaLaunchKernelEx(&cfg, link_kernel, mode, in, out, n, wc, (unsigned)CTAS, cnt[c] + k, (const unsigned*)epoch, err));
                }
                bump_epoch<<<1, 1, 0, st[c]>>>(epoch);
                CK(cudaStreamEndCapture(st[c], &g));
                CK(cudaGraphInstantiate(&ge[c], g, 0));
            }
            for (int w = 0; w < 5; ++w) for (int c = 0; c < chains; ++c) CK(cudaGraphLaunch(ge[c], st[c]));
            CK(cudaDeviceSynchronize());
            cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            std::vector<cudaEvent_t> done(chains);
            CK(cudaEventRecord(e0, st[0]));
            for (int c = 1; c < chains; ++c) CK(cudaStreamWaitEvent(st[c], e0, 0));
            for (int r = 0; r < REPS; ++r) for (int c = 0; c < chains; ++c) CK(cudaGraphLaunch(ge[c], st[c]));
            for (int c = 1; c < chains; ++c) { CK(cudaEventCreate(&done[c])); CK(cudaEventRecord(done[c], st[c])); CK(cudaStreamWaitEvent(st[0], done[c], 0)); }
            CK(cudaEventRecord(e1, st[0]));
            CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            int herr = 0; CK(cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost));
            float v = 0; CK(cudaMemcpy(&v, buf[0] + ((NK & 1) ? n : 0), 4, cudaMemcpyDeviceToHost));
            printf("chains %d  ctas %2d  mode %-5s: %.2f us per link (%d links + 1 per replay; result %.0f, expected %d; timeout flag %d)\n",
                   chains, CTAS, mode == 0 ? "plain" : mode == 1 ? "pdl" : "flag", ms * 1e3 / REPS / (NK + 1), NK, v, NK * (REPS + 5), herr);
        }
    }
    return 0;
}
