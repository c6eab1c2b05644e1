// Measurement behind DESIGN.md 5 "streamed host step": what chunking the 84 MB copy-in costs, and what each way of
// signalling "this piece has landed" between the pieces adds (B200, PCIe 5 x16: 1.52 ms in one piece; 64 Ki-character
// pieces 1.64 ms; + cuStreamWriteValue32 1.72 ms; + 4-byte memset 1.79 ms; event records are free but a kernel cannot
// poll them).  Build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o copy_gap tools/copy_gap.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
typedef CUresult (*WriteFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
int main() {
    const size_t n = 1000000; const size_t b_tf = n * 40, b_ph = n * 44;
    char *h_tf, *h_ph, *d_tf, *d_ph; uint32_t * d_gate;
    cudaMallocHost(&h_tf, b_tf); cudaMallocHost(&h_ph, b_ph); cudaMalloc(&d_tf, b_tf); cudaMalloc(&d_ph, b_ph); cudaMalloc(&d_gate, 64);
    void * w = nullptr; cudaDriverEntryPointQueryResult r; cudaGetDriverEntryPoint("cuStreamWriteValue32", &w, cudaEnableDefault, &r);
    WriteFn wr = (WriteFn)w;
    cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    std::vector<cudaEvent_t> evs(64); for (auto & e : evs) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    for (int mode = 0; mode < 5; ++mode) {
        for (size_t piece : {size_t(1000000), size_t(262144), size_t(65536), size_t(32768)}) {
            float best = 1e9;
            for (int rep = 0; rep < 6; ++rep) {
                cudaEventRecord(e0, s);
                int k = 0;
                for (size_t first = 0; first < n; first += piece, ++k) {
                    size_t cnt = (n - first < piece) ? n - first : piece;
                    cudaMemcpyAsync(d_tf + first * 40, h_tf + first * 40, cnt * 40, cudaMemcpyHostToDevice, s);
                    cudaMemcpyAsync(d_ph + first * 44, h_ph + first * 44, cnt * 44, cudaMemcpyHostToDevice, s);
                    if (mode == 1) wr(s, (CUdeviceptr)d_gate, (unsigned)(first + cnt), 0);
                    if (mode == 2) cudaEventRecord(evs[k % 64], s);
                    if (mode == 3) cudaMemsetAsync(d_gate, k, 4, s);
                    if (mode == 4) wr(s, (CUdeviceptr)d_gate, (unsigned)(first + cnt), CU_STREAM_WRITE_VALUE_NO_MEMORY_BARRIER);
                }
                cudaEventRecord(e1, s); cudaStreamSynchronize(s);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            printf("mode %d (%s) piece %7zu: %.3f ms\n", mode, mode == 0 ? "copies only" : mode == 1 ? "write-value" : mode == 2 ? "event record" : mode == 3 ? "memset 4B" : "write-value no-barrier", piece, best);
        }
    }
    return 0;
}
