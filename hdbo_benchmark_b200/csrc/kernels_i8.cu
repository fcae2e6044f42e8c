// Exact digit-sliced ("Ozaki"-style) GEMMs of the posterior on the 5th-generation tensor cores.
//
// tcgen05 has no f64 kind, and the FP64 DMMA kernels (kernels_gemm.cu) already run at 95 % of the 37 TFLOP/s DMMA ceiling.  The two
// GEMMs of the no-grad pool evaluation are therefore ALSO available as exact integer computations on tcgen05.mma kind::i8:
//   k_cross_i8   G = Zs Xs^T (scaled candidates x scaled training points, K = d)   ->  r2 -> k(r2) -> digits of K*, mean partials
//   k_var_i8     V = K* L^-T (K = n, triangular)                                   ->  row sums of squares (variance partials)
// Both operands are split into 6 base-256 digit planes (i8.cuh); the 21 plane pairs with s + t <= 5 are exact int8 GEMMs with
// int32 accumulation, the 6 weight groups g = s + t live side by side in TMEM (6 x 64 columns), and only the final recombination
// sum_g 2^-(8g+c) acc_g is rounded (FP64).  Measured against the FP64 path: variance within 1e-12 of the prior variance, mean
// within 1e-11 relative (tests/test_gpu_i8.py).
//
// Kernel shape (both): persistent, one CTA per SM walking a static list of (128 candidates x 64 columns) tiles; warp 0 = producer
// (two 1-D bulk copies per 32-byte k-block bring all 6+6 digit planes into a ring that keeps running across tiles), warp 1 = MMA
// issuer (8 merged MMAs per k-block), remaining warps = epilogue from TMEM, which hand TMEM back to the issuer as soon as they have
// drained it.  SASS: UTCIMMA / UBLKCP / LDTM / UTCBAR.
#include "kernels.cuh"
#include "i8.cuh"
#include "tc05.cuh"

namespace hdbo {

// Operand layout ("tile images").  Digit operands are stored in global memory exactly as the tensor core wants them in shared
// memory, so one 1-D bulk copy (cp.async.bulk, full 128-byte lines) per operand and k-block fills a pipeline stage:
//   image(tile, kb) = [S planes][ROWS rows][32 bytes of k], SWIZZLE_32B applied (16-byte halves of rows 4..7 of every 8 swapped),
//   stored at ((tile * nkb + kb) * S * ROWS * 32);  ROWS = 128 for the A side (candidates), 64 for the B side.
// (A 3-D tensor map over plain [S][rows][K] planes works too -- first version -- but its 32-byte box rows are one L2 request each
// and the L1->XBAR request path saturated at 75 %: profiles/r01_ncu_full_i8_v1.raw.csv.)

// position of k-byte `lane` (0..31) of row r inside its 32-byte image row
__device__ __forceinline__ int img_inrow(int r, int lane) { return (((lane >> 4) ^ ((r >> 2) & 1)) << 4) | (lane & 15); }

// digit images of L^-1 (once per fit): one warp per row.  scale[i] = 2^e_i
__global__ void k_slice_linv(const double* __restrict__ Linv, int ld, int n_pad, int8_t* __restrict__ img, double* __restrict__ scale) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n_pad) return;
    const size_t b = blockIdx.y;                         // batch: matrices / image sets / scale vectors are contiguous per GP
    Linv += b * (size_t)n_pad * ld; img += b * (size_t)I8_S * n_pad * n_pad; scale += b * n_pad;
    const double* src = Linv + (size_t)row * ld;
    double mx = 0.0;
    for (int j = lane; j <= row; j += 32) mx = fmax(mx, fabs(src[j]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const int ex = i8_row_exponent(mx);
    const double down = ldexp(1.0, I8_SIGNED_FRAC_BITS - ex);
    if (lane == 0) scale[row] = ldexp(1.0, ex);
    const int nkb = n_pad >> 5, bn = row >> 6, r = row & 63;
    const int inrow = img_inrow(r, lane);
    for (int kb = 0; kb < nkb; kb++) {
        const int j = kb * 32 + lane;
        const long long x = (j <= row) ? __double2ll_rn(src[j] * down) : 0LL;
        unsigned dg[I8_S];
        i8_signed_digits(x, dg);
        int8_t* dst = img + ((size_t)bn * nkb + kb) * I8_B_IMAGE + r * 32 + inrow;
#pragma unroll
        for (int t = 0; t < I8_S; t++) dst[t * 64 * 32] = (int8_t)dg[t];
    }
}

cudaError_t launch_slice_linv(const double* Linv, int ld, int n_pad, int8_t* img, double* scale, int batch, cudaStream_t st) {
    k_slice_linv<<<dim3((n_pad + 7) / 8, batch), 256, 0, st>>>(Linv, ld, n_pad, img, scale);
    return cudaGetLastError();
}

// Digit images of centred, lengthscale-scaled input rows (candidates: ROWS = 128 per tile, the A side; training points:
// ROWS = 64, the B side).  One warp per row: zs_j = (x_j - c_j) / l_j, row scale 2^e, digits of zs 2^-e; the k-range is
// padded with zeros to a multiple of 32.  norms (optional) = sum_j zs_j^2 in FP64.
template <int ROWS>
__global__ void k_slice_rows(const double* __restrict__ X, long long ldx, int rows_valid, int rows_pad, int d,
                             const double* __restrict__ center, const double* __restrict__ inv_ls, int8_t* __restrict__ img,
                             double* __restrict__ scale, double* __restrict__ norms, long long bs_rows) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows_pad) return;
    const int nkb = (d + 31) >> 5, tile = row / ROWS, r = row % ROWS;
    const size_t b = blockIdx.y;                         // batch: per-GP lengthscales; outputs strided by bs_rows rows
    inv_ls += b * d; img += b * (size_t)bs_rows * nkb * 32 * I8_S; scale += b * bs_rows;
    if (norms) norms += b * bs_rows;
    const bool live = row < rows_valid;
    const double* src = X + (size_t)row * ldx;
    double mx = 0.0, ss = 0.0;
    if (live)
        for (int j = lane; j < d; j += 32) { const double z = (src[j] - center[j]) * inv_ls[j]; mx = fmax(mx, fabs(z)); ss = fma(z, z, ss); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o)); ss += __shfl_xor_sync(0xffffffffu, ss, o); }
    const int ex = i8_row_exponent(mx);
    const double down = ldexp(1.0, I8_SIGNED_FRAC_BITS - ex);
    if (lane == 0) { scale[row] = ldexp(1.0, ex); if (norms) norms[row] = ss; }
    const int inrow = img_inrow(r, lane);
    for (int kb = 0; kb < nkb; kb++) {
        const int j = kb * 32 + lane;
        long long x = 0;
        if (live && j < d) x = __double2ll_rn((src[j] - center[j]) * inv_ls[j] * down);
        unsigned dg[I8_S];
        i8_signed_digits(x, dg);
        int8_t* dst = img + ((size_t)tile * nkb + kb) * (I8_S * ROWS * 32) + r * 32 + inrow;
#pragma unroll
        for (int t = 0; t < I8_S; t++) dst[t * ROWS * 32] = (int8_t)dg[t];
    }
}

cudaError_t launch_slice_rows(int rows_per_tile, const double* X, long long ldx, int rows_valid, int rows_pad, int d, const double* center,
                              const double* inv_ls, int8_t* img, double* scale, double* norms, long long bs_rows, int batch,
                              cudaStream_t st) {
    if (rows_pad == 0) return cudaSuccess;
    const dim3 grid((rows_pad + 7) / 8, batch);
    if (rows_per_tile == 128) k_slice_rows<128><<<grid, 256, 0, st>>>(X, ldx, rows_valid, rows_pad, d, center, inv_ls, img, scale, norms, bs_rows);
    else k_slice_rows<64><<<grid, 256, 0, st>>>(X, ldx, rows_valid, rows_pad, d, center, inv_ls, img, scale, norms, bs_rows);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------------
namespace {
constexpr int VT_M = 128, VT_N = 64, VT_KB = 32;                  // k-block = one 32-byte swizzle row = one MMA k-step
constexpr int VT_STAGES = 5;
constexpr int VT_A_STAGE = I8_S * VT_M * VT_KB;                   // 24576
constexpr int VT_B_STAGE = I8_S * VT_N * VT_KB;                   // 12288
constexpr int VT_GN = 16;   // rasterisation of k_var_i8: a wave of 148 CTAs covers ~9 row tiles x 16 column tiles
constexpr int VT_SMEM = VT_STAGES * (VT_A_STAGE + VT_B_STAGE) + 256 + 1024;
static_assert(VT_A_STAGE == I8_A_IMAGE && VT_B_STAGE == I8_B_IMAGE, "image sizes");
static_assert(I8_S * VT_N + 2 * I8_S * 8 <= 512, "weight groups + double-buffered A planes must fit the 512 TMEM columns");
#ifndef HDBO_VAR_A_TMEM
#define HDBO_VAR_A_TMEM false   // measured: A planes via tcgen05.cp + TS-mode MMAs are exact but 20 % SLOWER (31.6 -> 38.1 ms per 2^18 pool)
#endif

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// Pipeline shared by k_var_i8 and k_cross_i8.  PERSISTENT kernels: one CTA per SM walks a static list of tiles
// (t = blockIdx.x, + gridDim.x, ...); warp 0 = producer, warp 1 = MMA issuer, warps >= 2 = epilogue.  The operand ring keeps
// running across tile boundaries (the producer prefetches the next tile's k-blocks while the epilogue of the current one runs), the
// MMA issuer starts the next tile as soon as the epilogue warps have drained TMEM (`tmem_empty`), so TMEM allocation, barrier
// set-up, pipeline fill and most of the epilogue are off the critical path.
struct I8Pipe {
    uint8_t *As, *Bs, *extra;     // extra: the bytes after the barriers (kernel-specific staging)
    uint32_t full0, empty0, accum, tmem_empty, tmem_base;
};

template <int STAGES>
__device__ __forceinline__ I8Pipe i8_pipe_setup(uint8_t* smem_raw, int epilogue_warps) {
    uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    I8Pipe pp;
    pp.As = base;
    pp.Bs = base + STAGES * VT_A_STAGE;
    uint64_t* bars = (uint64_t*)(base + STAGES * (VT_A_STAGE + VT_B_STAGE));
    pp.extra = base + STAGES * (VT_A_STAGE + VT_B_STAGE) + 256;
    pp.full0 = smem_u32(bars); pp.empty0 = smem_u32(bars + STAGES); pp.accum = smem_u32(bars + 2 * STAGES);
    pp.tmem_empty = smem_u32(bars + 2 * STAGES + 1);
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * STAGES + 2);
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(pp.full0 + 8 * s, 1); mbar_init(pp.empty0 + 8 * s, 1); }
        mbar_init(pp.accum, 1);
        mbar_init(pp.tmem_empty, epilogue_warps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if ((threadIdx.x >> 5) == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    pp.tmem_base = *tmem_slot;
    return pp;
}

__device__ __forceinline__ void i8_pipe_teardown(const I8Pipe& pp) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if ((threadIdx.x >> 5) == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(pp.tmem_base), "r"(512) : "memory");
}

// producer (one thread): two bulk copies per k-block, images asrc[kb], bsrc[kb]; kbg = running k-block count of this CTA
template <int STAGES>
__device__ __forceinline__ bool i8_pipe_produce(const I8Pipe& pp, const int8_t* asrc, const int8_t* bsrc, int kblocks, uint32_t& kbg, int* flag) {
    for (int kb = 0; kb < kblocks; kb++, kbg++) {
        const uint32_t s = kbg % STAGES, ph = (kbg / STAGES) & 1;
        if (!mbar_wait(pp.empty0 + 8 * s, ph ^ 1, flag)) return false;
        mbar_expect_tx(pp.full0 + 8 * s, VT_A_STAGE + VT_B_STAGE);
        bulk_load(smem_u32(pp.As + s * VT_A_STAGE), asrc + (size_t)kb * VT_A_STAGE, VT_A_STAGE, pp.full0 + 8 * s);
        bulk_load(smem_u32(pp.Bs + s * VT_B_STAGE), bsrc + (size_t)kb * VT_B_STAGE, VT_B_STAGE, pp.full0 + 8 * s);
    }
    return true;
}

// MMA issuer (one thread), one tile.  The B planes t0..t0+c-1 of a stage are contiguous (64 rows x 32 B each), i.e. ONE K-major
// operand with N = 64 c rows, and the accumulators of the weight groups g = s + t are adjacent 64-column blocks of TMEM: digit
// plane s of A against planes t0.. of B is a single MMA with N = 64 c writing groups s+t0...  8 MMAs (N <= 256) instead of 21 per
// k-block: the 4 KB A plane is read from shared memory 8 instead of 21 times (SS-mode operand reads were the limiter of the
// unmerged version: 192 B/clk needed for N = 64 against 128 B/clk of shared-memory bandwidth).
// `tile_it` = how many tiles this CTA has finished: the accumulators may be overwritten once the epilogue has drained them.
template <int STAGES, bool A_SIGNED, bool A_TMEM = false>
__device__ __forceinline__ bool i8_pipe_mma(const I8Pipe& pp, int kblocks, uint32_t& kbg, uint32_t tile_it, int* flag) {
    if (tile_it > 0 && !mbar_wait(pp.tmem_empty, (tile_it - 1) & 1, flag)) return false;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int kb = 0; kb < kblocks; kb++, kbg++) {
        const uint32_t s = kbg % STAGES, ph = (kbg / STAGES) & 1;
        if (!mbar_wait(pp.full0 + 8 * s, ph, flag)) return false;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a0 = smem_u32(pp.As + s * VT_A_STAGE), b0 = smem_u32(pp.Bs + s * VT_B_STAGE);
        const uint32_t acc = kb != 0;
        // A_TMEM: the 6 A planes of this k-block are first copied to TMEM (columns 384.., double-buffered by k-block parity; the
        // copy and the MMAs execute in issue order), so the MMAs read only B from shared memory.
        const uint32_t at0 = pp.tmem_base + I8_S * VT_N + (kbg & 1) * (I8_S * 8);
        if (A_TMEM) {
#pragma unroll
            for (int sa = 0; sa < I8_S; sa++) utccp_128x256b(at0 + sa * 8, make_desc_sw32(a0 + sa * VT_M * VT_KB));
        }
#define HDBO_I8_MMA(SA, T0, CNT, ACC)                                                                                         \
        if (A_TMEM) umma_i8_ts(pp.tmem_base + ((SA) + (T0)) * VT_N, at0 + (SA) * 8, make_desc_sw32(b0 + (T0) * VT_N * VT_KB), \
                               idesc_i8((CNT) * VT_N, A_SIGNED), (ACC));                                                      \
        else umma_i8(pp.tmem_base + ((SA) + (T0)) * VT_N, make_desc_sw32(a0 + (SA) * VT_M * VT_KB),                            \
                     make_desc_sw32(b0 + (T0) * VT_N * VT_KB), idesc_i8((CNT) * VT_N, A_SIGNED), (ACC))
        HDBO_I8_MMA(0, 0, 4, acc); HDBO_I8_MMA(0, 4, 2, acc);      // plane 0 of A initialises all 6 groups at kb = 0
        HDBO_I8_MMA(1, 0, 4, 1u);  HDBO_I8_MMA(1, 4, 1, 1u);
        HDBO_I8_MMA(2, 0, 4, 1u);
        HDBO_I8_MMA(3, 0, 3, 1u);
        HDBO_I8_MMA(4, 0, 2, 1u);
        HDBO_I8_MMA(5, 0, 1, 1u);
#undef HDBO_I8_MMA
        umma_commit(pp.empty0 + 8 * s);
    }
    umma_commit(pp.accum);
    return true;
}

// epilogue warps: the accumulators of this CTA's tile number `tile_it` are complete
__device__ __forceinline__ bool i8_pipe_wait_accum(const I8Pipe& pp, uint32_t tile_it, int* flag) {
    const bool ok = mbar_wait(pp.accum, tile_it & 1, flag);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    return ok;
}
// epilogue warps (all lanes): this warp has read everything it needs from TMEM
__device__ __forceinline__ void i8_pipe_release_tmem(const I8Pipe& pp) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    if ((threadIdx.x & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(pp.tmem_empty) : "memory");
}

static int i8_sm_count() {   // SM count of the calling thread's CURRENT device (a process may drive several GPUs)
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    return n;
}

// ---------------------------------------------------------------------------------------------------------------------
// K*(Z, X) on the int8 tensor cores: the scaled-candidate x scaled-train inner products as 21 exact digit-pair GEMMs
// (K = roundup(d, 32)), then per element  r2 = |z|^2 + |x|^2 - 2 z.x  ->  k(r2)  ->  the 6 bytes of rint(k 2^48) (the A operand
// of k_var_i8, written straight into its tile images: a thread owns one candidate row, i.e. whole 16-byte halves of image rows)
// and the partial posterior mean over this tile's 64 training points.
struct CrossI8Params {
    const int8_t* Zimg;       // candidate digit images (tile = 128 rows), nkb_d k-blocks
    const int8_t* Ximg;       // training-point digit images (tile = 64 rows)
    const double *zscale, *xscale, *zn, *xn, *alpha, *hyp;
    int8_t* Kimg;             // out: K*/outputscale digit images (tile = 128 rows, n_pad / 32 k-blocks)
    double* meanpart;         // out: [ntiles64][rows_pad]
    int m, n, rows_pad, ntiles64, ntiles, nkb_d, nkb_n;
    int* flag;
    // batch (SAAS hyper-samples): tiles [b][bm][bn]; per-GP strides
    int tiles_per_gp, n_pad, rows_alloc;
    long long bsK;            // bytes between the K* image sets of consecutive GPs
};

// 576 threads: warp 0 producer, warp 1 MMA issuer, warps 2-17 epilogue (four warps per TMEM lane quarter, 16 columns each).
//
// Measured on B200 (tools/fp64_vs_tc.cu, profiles/r01_fp64_vs_tensor.jsonl): the scalar FP64 pipe is starved while tcgen05.mma is
// streaming on the same SM (a DFMA stream runs 4-14x slower next to a saturated int8 MMA stream; FFMA and integer work do not
// care).  The MMAs of the CTA's next tile therefore do NOT hide the FP64 part of this epilogue -- a tile costs MMA time + FP64 time --
// and every FP64 instruction saved is tile time saved.  The epilogue is written for a minimal FP64 count (26 per K* element; the
// first version needed 44 plus 13 conversion-pipe instructions):
//   * the 6 accumulator groups are combined as two 48-bit integers (3 IMAD.WIDE each), dropped into the mantissa of 1.5 2^52 and
//     recovered with one FP64 subtraction each -- no I2F;
//   * x = KF r^2 = KF(|z|^2 + |x|^2) - 2 KF 2^(ez + ex) G with the power-of-two column scale ADDED TO THE EXPONENT FIELD of the row
//     constant (integer add), KF = 5 for Matern-5/2 folded into the norms so that a = sqrt(x), k = (1 + a + x/3) exp(-a);
//   * clamps are integer min / max on the high word; the floor x0 is where k = 1 - 2^-48 (the largest digit string: what a zero
//     distance gave before, too), the ceiling keeps exp() out of the underflow range, so neither needs a special case; padding
//     rows / columns get a huge norm instead of a select (k < 1e-270: zero digits, no visible mean contribution);
//   * sqrt: 22-bit seed + one cubic correction (5 FP64), exp: 64-entry table + degree-5 polynomial (10 FP64)  (fast_math.cuh);
//   * outputscale folded into alpha; rint(k 2^48) read from the mantissa of k 2^48 + 2^52.
// Order inside a tile: recombination -> the whole FP64 part (tensor pipe idle, TMEM still held) -> TMEM release -> digit gathering,
// staging, barriers and the bulk store, which overlap the next tile's MMAs (releasing TMEM right after the recombination let the
// FP64 part crawl through the MMA phase instead: 8.6 -> 8.0 ms per 2^18 candidates).
constexpr int CE_WARPS = 16, CE_THREADS = 64 + CE_WARPS * 32;
constexpr int CE_STAGES = 4;                       // 4-stage operand ring + 48 KB staging of the two output images of a tile
constexpr int CE_SMEM = CE_STAGES * (VT_A_STAGE + VT_B_STAGE) + 256 + 2 * VT_A_STAGE + 1024;

// 16 columns of the 6 weight groups (group g at TMEM column g * VT_N, weight 2^-(8g + W0)) -> 16 doubles.  Groups 3h .. 3h+2 carry
// relative weights 1, 2^-8, 2^-16: summed as a 48-bit integer (|sum| < 2^48 because |group| < 2^31), converted through the
// mantissa of 1.5 2^52.  4 FP64 instructions per value (the I2F / DFMA form needs 6 + 6 conversions).
template <int W0>
__device__ __forceinline__ void i8_recombine16(uint32_t taddr, double (&v)[16]) {
#pragma unroll
    for (int h = 1; h >= 0; h--) {
        uint32_t r0[16], r1[16], r2[16];
        tmem_ld16_nowait(taddr + (3 * h) * VT_N, r0);
        tmem_ld16_nowait(taddr + (3 * h + 1) * VT_N, r1);
        tmem_ld16_nowait(taddr + (3 * h + 2) * VT_N, r2);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const long long s48 = mad_wide((int32_t)r0[j], 65536, mad_wide((int32_t)r1[j], 256, mad_wide((int32_t)r2[j], 1, 0x4338000000000000LL)));
            const double dv = __longlong_as_double(s48) - 6755399441055744.0;
            v[j] = h ? dv * i8_pow2(-(40 + W0)) : fma(dv, i8_pow2(-(16 + W0)), v[j]);
        }
    }
}

template <int KIND>
__global__ void __launch_bounds__(CE_THREADS, 1) k_cross_i8(CrossI8Params p) {   // (18 warps are allocated as 20: 96 registers is the cap)
    extern __shared__ uint8_t smem_raw[];
    __shared__ double s_xn[VT_N], s_al[VT_N], s_ms[3][VT_M], s_tab[64];
    __shared__ int s_xe[VT_N];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr double KF = (KIND == 0) ? 5.0 : 1.0; // Matern-5/2 works on x = 5 r^2
    if (tid < 64) s_tab[tid] = EXP2_TAB64[tid];    // (the barrier inside i8_pipe_setup publishes it)
    const I8Pipe pp = i8_pipe_setup<CE_STAGES>(smem_raw, CE_WARPS);
    uint8_t* stage_out = pp.extra;                 // [2 k-blocks][6 planes][128 rows][32 B]: the tile's K* digit images, as in global memory
    if (warp == 0 && lane == 0) {
        uint32_t kbg = 0;
        for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
            const int b = t / p.tiles_per_gp, tt = t - b * p.tiles_per_gp, bn = tt % p.ntiles64, bm = tt / p.ntiles64;
            const int8_t* zi = p.Zimg + (size_t)b * p.rows_alloc * p.nkb_d * 32 * I8_S;
            const int8_t* xi = p.Ximg + (size_t)b * p.n_pad * p.nkb_d * 32 * I8_S;
            if (!i8_pipe_produce<CE_STAGES>(pp, zi + (size_t)bm * p.nkb_d * VT_A_STAGE, xi + (size_t)bn * p.nkb_d * VT_B_STAGE, p.nkb_d, kbg, p.flag)) break;
        }
    } else if (warp == 1 && lane == 0) {
        uint32_t kbg = 0, it = 0;
        for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, it++)
            if (!i8_pipe_mma<CE_STAGES, true>(pp, p.nkb_d, kbg, it, p.flag)) break;
    }
    __syncwarp();
    if (warp >= 2) {
        const int q = warp & 3, sl = (warp - 2) >> 2, row = q * 32 + lane;   // sl: 16-column slice of the tile
        const int swz = (row >> 2) & 1;
        const uint32_t stage_dst = smem_u32(stage_out) + (sl >> 1) * VT_A_STAGE + row * 32 + (((sl & 1) ^ swz) << 4);
        uint32_t it = 0;
        for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, it++) {
            const int b = t / p.tiles_per_gp, tt = t - b * p.tiles_per_gp, bn = tt % p.ntiles64, bm = tt / p.ntiles64, grow = bm * VT_M + row;
            if (tid >= 64 && tid < 64 + VT_N) {            // (the previous tile's readers passed the s_ms barrier below)
                const size_t c = (size_t)b * p.n_pad + bn * VT_N + tid - 64;
                const bool cv = bn * VT_N + tid - 64 < p.n;
                s_xn[tid - 64] = cv ? KF * p.xn[c] : 1e300;
                s_xe[tid - 64] = __double2hiint(p.xscale[c]) - 0x3FF00000;
                s_al[tid - 64] = cv ? p.hyp[b * 4] * p.alpha[c] : 0.0;
            }
            const size_t zr = (size_t)b * p.rows_alloc + grow;
            const double znr = (grow < p.m) ? KF * p.zn[zr] : 1e300, m2sz = -2.0 * KF * p.zscale[zr];
            const bool ok = i8_pipe_wait_accum(pp, it, p.flag);
            double v[16];                                  // G = (z / 2^ez) . (x / 2^ex); signed x signed digits: group weights 2^-(8g + 14)
            i8_recombine16<14>(pp.tmem_base + ((uint32_t)(q * 32) << 16) + sl * 16, v);
            if (tid == 64) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous tile's images have left the staging buffer
            asm volatile("bar.sync 1, %0;" ::"n"(CE_WARPS * 32) : "memory");   // s_xn / s_xe / s_al of this tile are in place
            // FP64 part first, with the tensor pipe idle (TMEM still held): v[j] <- k 2^48 + 2^52, whose mantissa is rint(k 2^48)
            double msum = 0.0;
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const int cl = sl * 16 + j;
                const double c = __hiloint2double(__double2hiint(m2sz) + s_xe[cl], __double2loint(m2sz));
                const double xr = fma(v[j], c, znr + s_xn[cl]);
                constexpr int HI_MIN = (KIND == 0) ? 0x3D180000 : 0x3D000000;   // 6 2^-48 (Matern: k = 1 - x/6), 2^-47 (RBF: k = 1 - x/2)
                constexpr int HI_MAX = (KIND == 0) ? 0x41186A00 : 0x4095E000;   // 4e5 (a <= 633), 1400 (t >= -700)
                const double x = __hiloint2double(max(min(__double2hiint(xr), HI_MAX), HI_MIN), __double2loint(xr));
                double kq;
                if (KIND == 0) {
                    const double a = sqrt_pos_cubic(x);
                    kq = fma(x, 3.33333333333333314830e-01, 1.0 + a) * exp_neg_tab_inrange(-a, s_tab);
                } else {
                    kq = exp_neg_tab_inrange(-0.5 * x, s_tab);
                }
                msum = fma(kq, s_al[cl], msum);
                v[j] = fma(kq, 281474976710656.0, 4503599627370496.0);
            }
            // now the MMA issuer may start this CTA's next tile: what is left of this one (byte gathering, staging, barriers, the
            // bulk store, the next tile's scalars) is integer / memory work, which DOES overlap the MMAs
            i8_pipe_release_tmem(pp);
            uint32_t dg[I8_S][4];                          // 16 columns = one 16-byte half of an image row per digit plane
#pragma unroll
            for (int jq = 0; jq < 4; jq++) {               // 4 columns -> one 32-bit word of every digit plane
                unsigned lo[4], hi[4];
#pragma unroll
                for (int e = 0; e < 4; e++) { lo[e] = (unsigned)__double2loint(v[jq * 4 + e]); hi[e] = (unsigned)__double2hiint(v[jq * 4 + e]) & 0xFFFFu; }
                // byte b of the four values -> one word (3 PRMT each)
#define HDBO_GATHER(SRC, SEL) __byte_perm(__byte_perm(SRC[0], SRC[1], SEL), __byte_perm(SRC[2], SRC[3], SEL), 0x5410)
                dg[5][jq] = HDBO_GATHER(lo, 0x0040); dg[4][jq] = HDBO_GATHER(lo, 0x0051);
                dg[3][jq] = HDBO_GATHER(lo, 0x0062); dg[2][jq] = HDBO_GATHER(lo, 0x0073);
                dg[1][jq] = HDBO_GATHER(hi, 0x0040); dg[0][jq] = HDBO_GATHER(hi, 0x0051);
#undef HDBO_GATHER
            }
#pragma unroll
            for (int tt = 0; tt < I8_S; tt++)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_dst + tt * VT_M * 32), "r"(dg[tt][0]), "r"(dg[tt][1]), "r"(dg[tt][2]), "r"(dg[tt][3]) : "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staging writes -> visible to the bulk-copy engine
            // (the integer clamps swallow NaN: a candidate row with a NaN coordinate -- NaN norm -- gets a NaN mean, as on the FP64 path)
            if (!ok || znr != znr) msum = __longlong_as_double(0x7ff8000000000000LL);
            if (sl) s_ms[sl - 1][row] = msum;
            asm volatile("bar.sync 1, %0;" ::"n"(CE_WARPS * 32) : "memory");
            if (!sl) p.meanpart[(size_t)b * p.ntiles64 * p.rows_alloc + (size_t)bn * p.rows_pad + grow] = ((msum + s_ms[0][row]) + s_ms[1][row]) + s_ms[2][row];
            if (tid == 64) {                               // (after the barrier above: every warp's staging writes are done)
                int8_t* gdst = p.Kimg + (size_t)b * p.bsK + ((size_t)bm * p.nkb_n + bn * 2) * VT_A_STAGE;   // k-blocks 2 bn, 2 bn + 1: contiguous
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(stage_out)), "r"(2 * VT_A_STAGE) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        if (tid == 64) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all images written before the CTA exits
    }
    i8_pipe_teardown(pp);
}

cudaError_t launch_cross_i8(int kind, const int8_t* Zimg, const int8_t* Ximg, const double* zscale, const double* xscale, const double* zn,
                            const double* xn, const double* alpha, const double* hyp, int8_t* Kimg, double* meanpart, int m, int n,
                            int rows_pad, int rows_alloc, int n_pad, int d, int batch, int* flag, cudaStream_t st) {
    static PerDeviceOnce attr_once; int attr_dev;
    if (attr_once.pending(attr_dev)) {
        cudaError_t e = cudaFuncSetAttribute(k_cross_i8<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, CE_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_cross_i8<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, CE_SMEM);
        if (e != cudaSuccess) return e;
        attr_once.mark(attr_dev);
    }
    const int per_gp = (rows_pad / VT_M) * (n_pad / VT_N), ntiles = per_gp * batch;
    CrossI8Params p{Zimg, Ximg, zscale, xscale, zn, xn, alpha, hyp, Kimg, meanpart, m, n, rows_pad, n_pad / VT_N, ntiles, (d + 31) / 32,
                    n_pad / VT_KB, flag, per_gp, n_pad, rows_alloc, (long long)rows_alloc * n_pad * 8};
    const int grid = ntiles < i8_sm_count() ? ntiles : i8_sm_count();
    if (kind == 0) k_cross_i8<0><<<grid, CE_THREADS, CE_SMEM, st>>>(p);
    else k_cross_i8<1><<<grid, CE_THREADS, CE_SMEM, st>>>(p);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------------
// V = K* L^-T and its row sums of squares.  Triangular k-range (L^-1 is lower triangular: only
// k-blocks <= column tile).
struct VarI8Params {
    const int8_t* Aimg;       // K* digit images (unsigned bytes), tile = 128 candidates
    const int8_t* Bimg;       // L^-1 digit images (balanced), tile = 64 rows of L^-1 (= columns of V)
    const double* colscale;   // [n_pad] 2^e_i of the L^-1 row = column of V
    const double* hyp;        // outputscale at [0]
    double* varpart;          // [ntiles64][rows_pad]
    int rows_pad, ntiles64, mtiles, nkb, nslots, nwaves;
    int* flag;                // set when a barrier wait timed out (results are then poisoned with NaN)
    int slots_per_gp, n_pad, rows_alloc;   // batch: slots [b][...]; per-GP strides derive from these
    long long bsK;                         // bytes between the K* image sets of consecutive GPs
};

// Tile list of k_var_i8.  Slots t are rasterised so that tiles processed at the same time by the 148 CTAs (t .. t+147) cover
// VT_GN column tiles x ~9 row tiles: every A (K* digit) tile is fetched from DRAM once per VT_GN column tiles instead of once per
// column tile (the A images of a chunk are 4x the L2).  Column-tile groups with the longest k-range come first.  CTA c takes the
// slots w G + c of the even "waves" w and w G + (G-1-c) of the odd ones (serpentine): with round-robin, CTA 0 always got the
// heaviest tile of a wave (4.2 % static imbalance in the k-block count, 7 % measured); serpentine balances it to < 0.1 %.
__device__ __forceinline__ bool var_tile(const VarI8Params& p, int wave, int& b, int& bm, int& bn) {
    const int G = gridDim.x, c = (wave & 1) ? G - 1 - (int)blockIdx.x : (int)blockIdx.x;
    int t = wave * G + c;
    if (t >= p.nslots) return false;
    b = t / p.slots_per_gp; t -= b * p.slots_per_gp;
    const int per_group = p.mtiles * VT_GN;
    const int grp = t / per_group, rem = t - grp * per_group;
    bm = rem / VT_GN; bn = p.ntiles64 - 1 - (grp * VT_GN + rem % VT_GN);
    return bn >= 0;                                                 // (ntiles64 not a multiple of VT_GN: empty slots)
}

// 320 threads: producer warp, MMA warp, 8 epilogue warps (two per TMEM lane quarter, 32 columns each): the accumulators are
// drained -- and TMEM handed back to the MMA issuer for the CTA's next tile -- in half the time of a 4-warp epilogue.
__global__ void __launch_bounds__(320, 1) k_var_i8(VarI8Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ double s_part[VT_M];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const I8Pipe pp = i8_pipe_setup<VT_STAGES>(smem_raw, 8);
    int b, bm, bn;
    if (warp == 0 && lane == 0) {
        uint32_t kbg = 0;
        for (int wv = 0; wv < p.nwaves; wv++) {
            if (!var_tile(p, wv, b, bm, bn)) continue;
            if (!i8_pipe_produce<VT_STAGES>(pp, p.Aimg + (size_t)b * p.bsK + (size_t)bm * p.nkb * VT_A_STAGE,
                                 p.Bimg + (size_t)b * I8_S * p.n_pad * p.n_pad + (size_t)bn * p.nkb * VT_B_STAGE,
                                 (bn + 1) * (VT_N / VT_KB), kbg, p.flag)) break;
        }
    } else if (warp == 1 && lane == 0) {
        uint32_t kbg = 0, it = 0;
        for (int wv = 0; wv < p.nwaves; wv++) {
            if (!var_tile(p, wv, b, bm, bn)) continue;
            if (!i8_pipe_mma<VT_STAGES, false, HDBO_VAR_A_TMEM>(pp, (bn + 1) * (VT_N / VT_KB), kbg, it++, p.flag)) break;
        }
    }
    __syncwarp();
    if (warp >= 2) {                         // ---- epilogue: FP64 recombination of the 6 weight groups, column scale, row sum of squares ----
        const int q = warp & 3, ch = (warp - 2) >> 2, row = q * 32 + lane;   // TMEM lane quarter, 32-column half of the tile
        uint32_t it = 0;
        for (int wv = 0; wv < p.nwaves; wv++) {
            if (!var_tile(p, wv, b, bm, bn)) continue;
            const double os = p.hyp[b * 4];
            const bool ok = i8_pipe_wait_accum(pp, it++, p.flag);
            // unsigned x signed digits: group weights 2^-(8g + 15).  Integer recombination (i8_recombine16): the tensor pipe idles until
            // TMEM is released, and 192 I2F per thread on the 16-lane conversion pipe were most of that time
            double v[32];
            {
                double (&v0)[16] = *reinterpret_cast<double (*)[16]>(&v[0]);
                double (&v1)[16] = *reinterpret_cast<double (*)[16]>(&v[16]);
                const uint32_t ta = pp.tmem_base + ((uint32_t)(q * 32) << 16) + ch * 32;
                i8_recombine16<15>(ta, v0);
                i8_recombine16<15>(ta + 16, v1);
            }
            i8_pipe_release_tmem(pp);
            const double* cs = p.colscale + (size_t)b * p.n_pad + bn * VT_N + ch * 32;
            double ssq = 0.0;
#pragma unroll
            for (int j = 0; j < 32; j++) { const double x = v[j] * (os * cs[j]); ssq = fma(x, x, ssq); }
            if (!ok) ssq = __longlong_as_double(0x7ff8000000000000LL);
            if (ch) s_part[row] = ssq;
            asm volatile("bar.sync 1, 256;" ::: "memory");     // the 8 epilogue warps
            if (!ch) p.varpart[(size_t)b * p.ntiles64 * p.rows_alloc + (size_t)bn * p.rows_pad + bm * VT_M + row] = ssq + s_part[row];
            asm volatile("bar.sync 1, 256;" ::: "memory");     // s_part may be rewritten for the next tile
        }
    }
    i8_pipe_teardown(pp);
}

// Aimg: K* digit images for rows_pad / 128 row tiles; Bimg: L^-1 digit images (layout at the top of this file)
cudaError_t launch_var_i8(const int8_t* Aimg, int rows_pad, int rows_alloc, const int8_t* Bimg, const double* colscale, const double* hyp,
                          int n_pad, double* varpart, int batch, int* flag, cudaStream_t st) {
    static PerDeviceOnce attr_once; int attr_dev;
    if (attr_once.pending(attr_dev)) {
        cudaError_t e = cudaFuncSetAttribute(k_var_i8, cudaFuncAttributeMaxDynamicSharedMemorySize, VT_SMEM);
        if (e != cudaSuccess) return e;
        attr_once.mark(attr_dev);
    }
    const int mt = rows_pad / VT_M, nt = n_pad / VT_N, groups = (nt + VT_GN - 1) / VT_GN;
    const int per_gp = mt * VT_GN * groups, nslots = per_gp * batch;
    const long long live = (long long)mt * nt * batch;
    const int grid = live < i8_sm_count() ? (int)live : i8_sm_count();
    VarI8Params p{Aimg, Bimg, colscale, hyp, varpart, rows_pad, nt, mt, n_pad / VT_KB, nslots, (nslots + grid - 1) / grid, flag,
                  per_gp, n_pad, rows_alloc, (long long)rows_alloc * n_pad * 8};
    k_var_i8<<<grid, 320, VT_SMEM, st>>>(p);
    return cudaGetLastError();
}


}  // namespace hdbo
