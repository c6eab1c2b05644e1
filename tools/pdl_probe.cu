// pdl_probe.cu -- does programmatic dependent launch let frame f+1 of a persistent, claim-based kernel start in the warp
// slots that frame f's tail leaves empty?  (tools/: measurement helper, not part of the library)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pdl_probe tools/pdl_probe.cu && /tmp/pdl_probe
//
// A frame is `claims` work items of random duration (60-140 us of spinning) taken from a counter by 148 x 16 one-warp
// blocks.  Item g of frame f+1 may only start when item g of frame f is done (a flag per item).  Plain stream order
// serialises the frames; with the launch attribute, and every block calling griddepcontrol.launch_dependents at its start,
// the next frame's blocks are scheduled as this frame's blocks exit.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ unsigned long long now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

__global__ void __launch_bounds__(32, 16) k_frame(unsigned * counter, unsigned * counter_next, volatile unsigned * done, unsigned frame, unsigned claims,
                                                  const unsigned * dur_ns, unsigned long long * sink, int chained) {
    extern __shared__ unsigned char pad[];
    // the NEXT frame's counter is zeroed before this frame lets the next one in (its blocks start only after every block here
    // has passed launch_dependents)
    if (blockIdx.x == 0 && threadIdx.x == 0) { *counter_next = 0u; __threadfence(); }
    __syncwarp();
    if (chained) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    unsigned long long acc = 0;
    for (;;) {
        unsigned g = 0;
        if (threadIdx.x == 0) g = atomicAdd(counter, 1u);
        g = __shfl_sync(0xffffffffu, g, 0);
        if (g >= claims) break;
        if (threadIdx.x == 0) while (done[g] != frame) __nanosleep(100);   // the same item of the previous frame
        __syncwarp();
        const unsigned long long t0 = now();
        while (now() - t0 < dur_ns[(g * 7919u + frame * 104729u) % claims]) acc += pad[threadIdx.x];
        __threadfence();
        __syncwarp();
        if (threadIdx.x == 0) done[g] = frame + 1u;
    }
    if (acc == 0xdeadbeefull) *sink = acc;
}

int main() {
    const unsigned claims = 4630, frames = 40, blocks = 148 * 16;
    unsigned * counters, * done, * dur;
    unsigned long long * sink;
    CK(cudaMalloc(&counters, 4 * sizeof(unsigned)));
    CK(cudaMalloc(&done, claims * sizeof(unsigned)));
    CK(cudaMalloc(&dur, claims * sizeof(unsigned)));
    CK(cudaMalloc(&sink, 8));
    std::vector<unsigned> h(claims);
    srand(1);
    for (unsigned i = 0; i < claims; ++i) h[i] = 60000u + unsigned(rand() % 80000);
    CK(cudaMemcpy(dur, h.data(), claims * sizeof(unsigned), cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(k_frame, cudaFuncAttributeMaxDynamicSharedMemorySize, 13 * 1024));
    cudaStream_t s;
    CK(cudaStreamCreate(&s));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int chained = 0; chained < 2; ++chained) {
        CK(cudaMemset(counters, 0, 4 * sizeof(unsigned)));
        CK(cudaMemset(done, 0, claims * sizeof(unsigned)));
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0, s));
        for (unsigned f = 0; f < frames; ++f) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(32); cfg.dynamicSmemBytes = 13 * 1024; cfg.stream = s;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = chained ? 1 : 0;
            // four counters in rotation: frame f claims from f % 4 and zeroes (f + 1) % 4, last used by frame f - 3.  A block of frame f starts
            // only when every block of f - 1 is resident; the grids fill the machine, so f - 2 (and f - 3) are gone by then
            CK(cudaLaunchKernelEx(&cfg, k_frame, counters + f % 4, counters + (f + 1) % 4, (volatile unsigned *)done, f, claims, (const unsigned *)dur, sink, chained));
        }
        CK(cudaEventRecord(e1, s));
        CK(cudaStreamSynchronize(s));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        std::vector<unsigned> d(claims);
        CK(cudaMemcpy(d.data(), done, claims * sizeof(unsigned), cudaMemcpyDeviceToHost));
        unsigned bad = 0;
        for (unsigned i = 0; i < claims; ++i) bad += d[i] != frames;
        double work = 0;
        for (unsigned i = 0; i < claims; ++i) work += h[i];
        printf("%s: %.1f us per frame (work per slot %.1f us), %u items not at frame %u\n", chained ? "chained (PDL + item flags)" : "stream order",
               1e3 * ms / frames, work / blocks * 1e-3, bad, frames);
    }
    return 0;
}
