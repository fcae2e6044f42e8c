// A/B harness for the persistent rank-256 update kernel of the pipelined factorisation (k_side_updates, csrc/kernels_gemm.cu):
// C(i, j) -= A(i, 0..255) B(j, 0..255)^T over a strided job list, one CTA per SM, with globaltimer / clock64 stamps of one CTA's
// phases (C read, pipeline fill + mainloop, C write).  Variants: the CTA tile / k-block configuration, C added in the epilogue
// instead of read into the accumulators, an L2 prefetch of the next job's C tile, two independent 128x64 warp-group pipelines per
// CTA (staggered by half a job).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/side_job_bench tools/side_job_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../hdbo_benchmark_b200/csrc/gemm_core.cuh"
using namespace hdbo;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

struct Stamps { unsigned long long ns[64][4]; long long clk[64][4]; };

// job q -> C tile (i, j) of a T x T tile matrix, all distinct inside one launch; operands = tile rows i, j of Lm (K = kblocks * 16)
__device__ __forceinline__ void job_tiles(int q, int T, int& i, int& j) { i = q / T; j = q % T; }

template <class G, int EPI_ADD, int PREFETCH = 0>
__global__ void __launch_bounds__(G::THREADS, 1) k_side(const double* __restrict__ Lm, double* __restrict__ Cm, int ld, int T,
                                                        int njobs, int kblocks, Stamps* st) {
    extern __shared__ __align__(16) double smem[];
    const size_t lds = (size_t)ld;
    int n = 0;
    for (int job = blockIdx.x; job < njobs; job += gridDim.x, n++) {
        int i, j; job_tiles(job, T, i, j);
        const double* A = Lm + (size_t)i * 128 * lds;
        const double* B = Lm + (size_t)j * 128 * lds;
        double* C = Cm + (size_t)i * 128 * lds + (size_t)j * 128;
        const bool rec = st && blockIdx.x == 0 && threadIdx.x == 0 && n < 64;
        if (rec) { st->ns[n][0] = gtimer(); st->clk[n][0] = clock64(); }
        typename G::Frag acc;
        if (EPI_ADD) acc.zero();
        else {
#pragma unroll
            for (int mt = 0; mt < G::MT; mt++)
#pragma unroll
                for (int nt = 0; nt < G::NT; nt++) {
                    const double2 c = *reinterpret_cast<const double2*>(C + (size_t)G::frag_row(mt) * lds + G::frag_col(nt));
                    acc.v[mt][nt][0] = -c.x; acc.v[mt][nt][1] = -c.y;
                }
        }
        if (rec) { st->ns[n][1] = gtimer(); st->clk[n][1] = clock64(); }
        if (PREFETCH && job + (int)gridDim.x < njobs) {          // the next job's C tile -> L2 (1024 lines of 128 B, 4 per thread)
            int ni, nj; job_tiles(job + gridDim.x, T, ni, nj);
            const double* Cn = Cm + (size_t)ni * 128 * lds + (size_t)nj * 128;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int l = threadIdx.x + u * 256;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(Cn + (size_t)(l >> 3) * lds + (l & 7) * 16));
            }
        }
        G::mainloop(A, ld, B, ld, 0, kblocks * 16 / G::BK, smem, acc);
        if (rec) { st->ns[n][2] = gtimer(); st->clk[n][2] = clock64(); }
#pragma unroll
        for (int mt = 0; mt < G::MT; mt++)
#pragma unroll
            for (int nt = 0; nt < G::NT; nt++) {
                double* cp = C + (size_t)G::frag_row(mt) * lds + G::frag_col(nt);
                if (EPI_ADD) {
                    const double2 c = *reinterpret_cast<const double2*>(cp);
                    *reinterpret_cast<double2*>(cp) = make_double2(c.x - acc.v[mt][nt][0], c.y - acc.v[mt][nt][1]);
                } else {
                    *reinterpret_cast<double2*>(cp) = make_double2(-acc.v[mt][nt][0], -acc.v[mt][nt][1]);
                }
            }
        if (rec) { st->ns[n][3] = gtimer(); st->clk[n][3] = clock64(); }
    }
}

// two independent warp groups per CTA, each a GemmT<128, 64, 2, 2> pipeline (128 threads) on its own half-tile jobs; group 1 starts
// half a job late (a K/2 warm-up job on a scratch tile) so that one group's C traffic / pipeline fill falls into the other's mainloop
using GH = GemmT<128, 64, 2, 2>;
__global__ void __launch_bounds__(256, 1) k_side_2g(const double* __restrict__ Lm, double* __restrict__ Cm, int ld, int T, int njobs,
                                                    int kblocks, int stagger, Stamps* st) {
    extern __shared__ __align__(16) double smem_all[];
    const int grp = threadIdx.x >> 7;
    double* smem = smem_all + grp * (GH::SMEM_BYTES / 8);
    const size_t lds = (size_t)ld;
    int n = 0;
    if (grp == 1 && stagger) {               // warm-up: half a job of DMMAs on operands that are read anyway, result discarded
        GH::Frag acc; acc.zero();
        if (grp == 1) GH::mainloop<2>(Lm, ld, Lm, ld, 0, kblocks / 2, smem, acc);
        if (acc.v[0][0][0] == 123.456) Cm[0] = 0.0;
    }
    // half-tile job h of tile job q: columns [64 h, 64 h + 64) of C tile (i, j)
    for (int hj = 2 * blockIdx.x + grp; hj < 2 * njobs; hj += 2 * gridDim.x, n++) {
        int i, j; job_tiles(hj >> 1, T, i, j);
        const int h = hj & 1;
        const double* A = Lm + (size_t)i * 128 * lds;
        const double* B = Lm + ((size_t)j * 128 + 64 * h) * lds;
        double* C = Cm + (size_t)i * 128 * lds + (size_t)j * 128 + 64 * h;
        const bool rec = st && blockIdx.x == 0 && (threadIdx.x & 127) == 0 && grp == 0 && n < 64;
        if (rec) { st->ns[n][0] = gtimer(); st->clk[n][0] = clock64(); }
        GH::Frag acc;
#pragma unroll
        for (int mt = 0; mt < GH::MT; mt++)
#pragma unroll
            for (int nt = 0; nt < GH::NT; nt++) {
                const double2 c = *reinterpret_cast<const double2*>(C + (size_t)GH::frag_row(mt) * lds + GH::frag_col(nt));
                acc.v[mt][nt][0] = -c.x; acc.v[mt][nt][1] = -c.y;
            }
        if (rec) { st->ns[n][1] = gtimer(); st->clk[n][1] = clock64(); }
        if (grp == 0) GH::mainloop<1>(A, ld, B, ld, 0, kblocks, smem, acc);
        else GH::mainloop<2>(A, ld, B, ld, 0, kblocks, smem, acc);
        if (rec) { st->ns[n][2] = gtimer(); st->clk[n][2] = clock64(); }
#pragma unroll
        for (int mt = 0; mt < GH::MT; mt++)
#pragma unroll
            for (int nt = 0; nt < GH::NT; nt++)
                *reinterpret_cast<double2*>(C + (size_t)GH::frag_row(mt) * lds + GH::frag_col(nt)) =
                    make_double2(-acc.v[mt][nt][0], -acc.v[mt][nt][1]);
        if (rec) { st->ns[n][3] = gtimer(); st->clk[n][3] = clock64(); }
    }
}

__global__ void fill(double* p, size_t n, unsigned seed, double scale) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { unsigned x = (unsigned)(i * 2654435761u) ^ seed; x ^= x >> 13; x *= 0x5bd1e995u; x ^= x >> 15; p[i] = scale * ((double)(x & 0xffffff) / 16777216.0 - 0.5); }
}

static void report(const char* name, int ctas, int njobs, int kblocks, float ms, Stamps* dst) {
    Stamps h; CK(cudaMemcpy(&h, dst, sizeof(Stamps), cudaMemcpyDeviceToHost));
    const int rounds = (njobs + ctas - 1) / ctas;
    const int nrec = rounds < 64 ? rounds : 64;
    double rd = 0, ml = 0, wr = 0, tot = 0, clk = 0;
    for (int r = 1; r < nrec; r++) {          // skip the first (cold) job
        rd += (double)(h.ns[r][1] - h.ns[r][0]); ml += (double)(h.ns[r][2] - h.ns[r][1]); wr += (double)(h.ns[r][3] - h.ns[r][2]);
        tot += (double)(h.ns[r][3] - h.ns[r][0]);
        clk += (double)(h.clk[r][3] - h.clk[r][0]);
    }
    const int m = nrec > 1 ? nrec - 1 : 1;
    const double flops = 2.0 * 128 * 128 * kblocks * 16 * (double)njobs;
    printf("{\"variant\": \"%s\", \"ctas\": %d, \"jobs\": %d, \"K\": %d, \"launch_ms\": %.4f, \"us_per_round\": %.2f, \"tflops\": %.2f, "
           "\"cta0_us\": {\"c_read_issue\": %.2f, \"fill_mainloop\": %.2f, \"c_write_issue\": %.2f, \"job\": %.2f}, \"sm_mhz\": %.0f, "
           "\"pure_dmma_us_at_that_clock\": %.2f}\n",
           name, ctas, njobs, kblocks * 16, ms, ms * 1e3 / rounds, flops / ms / 1e9, rd / m / 1e3, ml / m / 1e3, wr / m / 1e3, tot / m / 1e3,
           clk / tot * 1e3, 128.0 * 128 * kblocks * 16 / 64.0 / (clk / tot * 1e3));
}

int main(int argc, char** argv) {
    const int ctas = argc > 1 ? atoi(argv[1]) : 132;
    const int rounds = argc > 2 ? atoi(argv[2]) : 4;
    const int kblocks = argc > 3 ? atoi(argv[3]) : 16;
    const int T = 32, ld = T * 128, njobs = ctas * rounds;
    if (njobs > T * T) { printf("too many jobs\n"); return 1; }
    double *Lm, *Cm; Stamps* st;
    CK(cudaMalloc(&Lm, sizeof(double) * ld * ld)); CK(cudaMalloc(&Cm, sizeof(double) * ld * ld)); CK(cudaMalloc(&st, sizeof(Stamps)));
    fill<<<(ld * ld + 255) / 256, 256>>>(Lm, (size_t)ld * ld, 1u, 1.0);
    fill<<<(ld * ld + 255) / 256, 256>>>(Cm, (size_t)ld * ld, 2u, 1.0);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto timeit = [&](auto launch) {
        float best = 1e30f;
        for (int r = 0; r < 8; r++) {
            CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (r > 1 && ms < best) best = ms;
        }
        return best;
    };
    CK(cudaFuncSetAttribute(k_side<G128, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, G128::SMEM_BYTES));
    CK(cudaFuncSetAttribute(k_side<G128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, G128::SMEM_BYTES));
    CK(cudaFuncSetAttribute(k_side<G128W, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, G128W::SMEM_BYTES));
    CK(cudaFuncSetAttribute(k_side_2g, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * GH::SMEM_BYTES));
    float ms;
    ms = timeit([&] { k_side<G128, 0><<<ctas, 256, G128::SMEM_BYTES>>>(Lm, Cm, ld, T, njobs, kblocks, st); });
    report("g128_bk16_c_into_acc", ctas, njobs, kblocks, ms, st);
    ms = timeit([&] { k_side<G128, 1><<<ctas, 256, G128::SMEM_BYTES>>>(Lm, Cm, ld, T, njobs, kblocks, st); });
    report("g128_bk16_c_in_epilogue", ctas, njobs, kblocks, ms, st);
    CK(cudaFuncSetAttribute(k_side<G128, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, G128::SMEM_BYTES));
    ms = timeit([&] { k_side<G128, 0, 1><<<ctas, 256, G128::SMEM_BYTES>>>(Lm, Cm, ld, T, njobs, kblocks, st); });
    report("g128_bk16_c_into_acc_l2_prefetch_next_c", ctas, njobs, kblocks, ms, st);
    ms = timeit([&] { k_side<G128W, 0><<<ctas, 256, G128W::SMEM_BYTES>>>(Lm, Cm, ld, T, njobs, kblocks, st); });
    report("g128_bk32_c_into_acc", ctas, njobs, kblocks, ms, st);
    ms = timeit([&] { k_side_2g<<<ctas, 256, 2 * GH::SMEM_BYTES>>>(Lm, Cm, ld, T, njobs, kblocks, 0, st); });
    report("2groups_128x64_lockstep", ctas, njobs, kblocks, ms, st);
    ms = timeit([&] { k_side_2g<<<ctas, 256, 2 * GH::SMEM_BYTES>>>(Lm, Cm, ld, T, njobs, kblocks, 1, st); });
    report("2groups_128x64_staggered", ctas, njobs, kblocks, ms, st);
    return 0;
}
