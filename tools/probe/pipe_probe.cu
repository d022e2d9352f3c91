// pipe_probe.cu -- issue rates of single instruction kinds and of ALU+FMA pipe mixes on sm_100a (developer tool).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/pipe_probe tools/probe/pipe_probe.cu ; run under gpurun.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ILP = 8, ITERS = 4096, THREADS = 512;

template <int WHICH>
__global__ void __launch_bounds__(THREADS) k(uint32_t *out, uint32_t seed, uint32_t one)
{
    uint32_t v[ILP], w[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { v[i] = seed * (threadIdx.x + 1) + i * 0x9E3779B9u; w[i] = v[i] ^ 0x55u; }
    uint32_t b = seed | 0x00010001u, c = (seed >> 3) | 1u;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (WHICH == 0) v[i] = __byte_perm(v[i], b, 0x4140 + i);                                   // PRMT
            if (WHICH == 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(b), "r"(c)); // IMAD
            if (WHICH == 2) v[i] = __umulhi(v[i], b) + c;                                              // IMAD.HI
            if (WHICH == 3) v[i] = __dp2a_lo(b, v[i], v[i]);                                           // IDP.2A
            if (WHICH == 4) v[i] = __dp4a(b, v[i], v[i]);                                              // IDP.4A
            if (WHICH == 5) { v[i] = __byte_perm(v[i], b, 0x4140 + i); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(w[i]) : "r"(b), "r"(c)); }   // PRMT + IMAD
            if (WHICH == 6) { v[i] = __byte_perm(v[i], b, 0x4140 + i); w[i] = __umulhi(w[i], b) + c; } // PRMT + IMAD.HI
            if (WHICH == 7) { v[i] = __byte_perm(v[i], b, 0x4140 + i); w[i] = __dp2a_lo(b, w[i], w[i]); }   // PRMT + IDP.2A
            if (WHICH == 8) { v[i] = __vabsdiffu4(v[i], b); w[i] = __fmaf_rn(__uint_as_float(w[i]), 1.0001f, 0.5f) > 0 ? w[i] + 1 : w[i]; }
            if (WHICH == 9) v[i] = __funnelshift_r(v[i], b, 8) ;                                       // SHF
            if (WHICH == 10) { v[i] = __byte_perm(v[i], b, 0x4140 + i); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(w[i]) : "r"(one), "r"(c)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(w[i]) : "r"(one), "r"(b)); }   // PRMT + 2 IMAD
        }
        b += 0x00010001u;
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) acc ^= v[i] ^ w[i];
    if (acc == 0x12345678u) out[blockIdx.x] = acc;
}

template <int WHICH>
static void run(const char *name, int nops, uint32_t *d, int blocks)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<WHICH><<<blocks, THREADS>>>(d, 12345u, 1u);
    cudaEventRecord(e0);
    k<WHICH><<<blocks, THREADS>>>(d, 6789u, 1u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double iters = (double)ITERS * ILP * THREADS * blocks;
    printf("%-18s %8.3f ms  %7.2f T iterations/s  (%d instr per iteration -> %7.2f T lane-instr/s)\n", name, ms, iters / ms / 1e9, nops,
           iters * nops / ms / 1e9);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 4;
    uint32_t *d;
    cudaMalloc(&d, blocks * 4);
    run<0>("PRMT", 1, d, blocks);
    run<1>("IMAD", 1, d, blocks);
    run<2>("IMAD.HI(+add)", 1, d, blocks);
    run<3>("IDP.2A", 1, d, blocks);
    run<4>("IDP.4A", 1, d, blocks);
    run<9>("SHF", 1, d, blocks);
    run<5>("PRMT+IMAD", 2, d, blocks);
    run<6>("PRMT+IMAD.HI", 2, d, blocks);
    run<7>("PRMT+IDP.2A", 2, d, blocks);
    run<10>("PRMT+2IMAD", 3, d, blocks);
    cudaFree(d);
    return 0;
}
