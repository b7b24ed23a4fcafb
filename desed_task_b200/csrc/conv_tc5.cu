// 3x3 convolution (Cin, Cout in {32, 64, 128}) on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), sm_100a.
//
// Implicit GEMM per CTA: D[128 pixels x Cout] (fp32 accumulator in TMEM, Cout columns) =
//   sum over 9 taps x Cin/32 channel chunks of  A_tap,chunk[128 px x 32 ch] * W_tap,chunk[Cout x 32 ci]^T   (kind::tf32).
// * A tiles come straight from the channels-last activation tensor through a 4-D TMA tensor map
//   {C, F, T, B} with box {32, TF, TT, 1}: the tap shift is a coordinate offset (f0+dx-1, t0+dy-1) and the zero padding of
//   the convolution is TMA's out-of-bounds zero fill - there is no im2col and no halo bookkeeping in the kernel.
// * W tiles come from the packed weights [tap][co][ci] through a 3-D map, box {32, Cout, 1}.
// * Both land in shared memory as K-major rows of 128 bytes with the 128-byte swizzle the UMMA descriptors expect.
// * Warp-specialised: warp 0 = TMA producer (one lane), warp 1 = TMEM allocator + MMA issuer (one lane issues
//   tcgen05.mma, tcgen05.commit releases pipeline stages through mbarriers), warps 2-5 = epilogue (tcgen05.ld of their
//   TMEM lane quadrant -> bias -> global store + per-channel sum / sum^2 for train-mode BatchNorm).
// The same kernel is the data-gradient when given the flipped/transposed weight pack.
#include "kernels.h"
#include "tc5.cuh"
#include <stdlib.h>

namespace sedk {
namespace {

constexpr int TC_STAGES = 6;
constexpr int TC_A_BYTES = 128 * 128;                 // 128 rows x 32 fp32
constexpr int TC_STAGE_BYTES = 2 * TC_A_BYTES;
constexpr int TC_THREADS = 192;
constexpr size_t TC_SMEM = (size_t)TC_STAGES * TC_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 1024 /*stats*/;

template <int CIN, int COUT, int TT, int TF>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv3x3_tc5_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const float* __restrict__ bias, float* __restrict__ out, double* __restrict__ stats, int T, int F,
                   int cmod) {
    pdl_enter();
    // cmod: number of REAL output channels behind the COUT MMA columns (COUT, or COUT / 2 in the paired-pixel mode where
    // column (h, co) is channel co of the pixel with parity h): bias and BatchNorm statistics are indexed modulo cmod
    static_assert(TT * TF == 128, "tile must hold 128 pixels");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;         // 1024-byte aligned (128B swizzle atoms)
    uint8_t* aligned = smem_raw + (base - smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(aligned + (size_t)TC_STAGES * TC_STAGE_BYTES);
    uint64_t* full = bars;                       // [TC_STAGES]
    uint64_t* empty = bars + TC_STAGES;          // [TC_STAGES]
    uint64_t* accum = bars + 2 * TC_STAGES;      // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 1);
    float* s_stat = reinterpret_cast<float*>(aligned + (size_t)TC_STAGES * TC_STAGE_BYTES + 256);   // [2][COUT]
    constexpr int TC_C = COUT;
    constexpr uint32_t TC_TMEM_COLS = COUT < 32 ? 32 : COUT;
    constexpr uint32_t TC_IDESC = tc_idesc(COUT);
    constexpr int NCHUNK = CIN / 32;
    constexpr uint32_t STAGE_TX = TC_A_BYTES + COUT * 128;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nTf = (F + TF - 1) / TF, nTt = (T + TT - 1) / TT;
    int tile = blockIdx.x;
    const int b = tile / (nTt * nTf);
    tile -= b * nTt * nTf;
    const int t0 = (tile / nTf) * TT, f0 = (tile % nTf) * TF;

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(accum, 1);
        fence_mbar_init();
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmB) : "memory");
    }
    for (int i = tid; i < 2 * TC_C; i += TC_THREADS) s_stat[i] = 0.f;
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                     "r"(TC_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    constexpr int NIT = 9 * NCHUNK;     // taps x 32-channel chunks
    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < NIT; it++) {
                const int s = it % TC_STAGES, ph = (it / TC_STAGES) & 1;
                mbar_wait_u32(smem_u32(&empty[s]), ph ^ 1);
                const int tap = it / NCHUNK, chunk = it - tap * NCHUNK;
                const int dy = tap / 3, dx = tap - dy * 3;
                const uint32_t a_dst = base + s * TC_STAGE_BYTES, b_dst = a_dst + TC_A_BYTES;
                mbar_expect_tx(&full[s], STAGE_TX);
                tma_load_4d(a_dst, &tmA, smem_u32(&full[s]), chunk * 32, f0 + dx - 1, t0 + dy - 1, b);
                tma_load_3d(b_dst, &tmB, smem_u32(&full[s]), chunk * 32, 0, tap);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int it = 0; it < NIT; it++) {
                const int s = it % TC_STAGES, ph = (it / TC_STAGES) & 1;
                mbar_wait_u32(smem_u32(&full[s]), ph);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint32_t a_src = base + s * TC_STAGE_BYTES, b_src = a_src + TC_A_BYTES;
#pragma unroll
                for (int k = 0; k < 4; k++) {       // UMMA_K = 8 tf32 = 32 bytes inside the 128-byte swizzle row
                    umma_tf32(tmem, umma_desc_sw128(a_src + k * 32), umma_desc_sw128(b_src + k * 32), TC_IDESC,
                              (it | k) != 0 ? 1u : 0u);
                }
                umma_commit(smem_u32(&empty[s]));   // frees the stage once these MMAs have read it
            }
            umma_commit(smem_u32(accum));           // accumulator complete
        }
    } else {
        // ---- epilogue: warp q owns TMEM lanes [32q, 32q+32) = accumulator rows (pixels)
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int ty = row / TF, tx = row - ty * TF;
        const int t = t0 + ty, f = f0 + tx;
        const bool valid = (t < T) && (f < F);
        mbar_wait_u32(smem_u32(accum), 0);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        float* orow = out + (((size_t)b * T + t) * F + f) * TC_C;
#pragma unroll 1
        for (int c = 0; c < COUT / 32; c++) {
            uint32_t v[32];
            tmem_ld32(v, tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32));
            float x[32];
#pragma unroll
            for (int j = 0; j < 32; j++) {
                x[j] = __uint_as_float(v[j]);
                if (bias != nullptr) x[j] += bias[(c * 32 + j) & (cmod - 1)];       // cmod is a power of two
            }
            if (valid) {
#pragma unroll
                for (int j = 0; j < 8; j++)
                    reinterpret_cast<float4*>(orow + c * 32)[j] = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
            }
            if (stats != nullptr) {
                float x2[32];
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    x[j] = valid ? x[j] : 0.f;
                    x2[j] = x[j] * x[j];
                }
                warp_transpose_reduce32(x, lane);
                warp_transpose_reduce32(x2, lane);
                atomicAdd(&s_stat[c * 32 + lane], x[0]);
                atomicAdd(&s_stat[TC_C + c * 32 + lane], x2[0]);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    }
    __syncthreads();
    if (stats != nullptr) {
        for (int i = tid; i < 2 * TC_C; i += TC_THREADS)
            atomicAdd(&stats[(i >= TC_C ? cmod : 0) + (i & (cmod - 1))], (double)s_stat[i]);
    }
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(TC_TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// Weight gradient on tcgen05:  dW[tap][co][ci] = sum_pixels gz[pix][co] * x[pix + tap][ci]   (Cout = 128, Cin in {64, 128}).
// UMMA with BOTH operands MN-major (the reduction index = pixels is the slow axis of the channels-last tensors; TF32
// MN-major operands use the 32-byte-atom flavour of the 128-byte swizzle on both the TMA and the descriptor side):
//   D[128 co x Cin] (+)= A^T B with A = gz tile [32 px x 128 co], B = shifted x tile [32 px x Cin], K = 8 pixels per MMA.
// One CTA owns one kernel row (3 taps: dx = 0,1,2 at a fixed dy) and keeps the 3 accumulators (3 x Cin TMEM columns) resident
// while it streams its share of the pixel tiles through a 3-stage TMA ring; gz is loaded once per stage and reused by the 3
// taps.  The epilogue adds the accumulators into the packed gradient with 16-byte vector atomics.
constexpr int WG_STAGES = 3;
constexpr int WG_KPIX = 32;                           // pixels per stage
constexpr int WG_CHUNK_BYTES = WG_KPIX * 128;         // one 32-channel chunk of one stage: 32 rows x 128 B
constexpr int WG_THREADS = 192;

// NDX = 3: convolution (CTA row dy = blockIdx.y, taps dx = 0..2);  NDX = 1: a single un-shifted tap, i.e. the plain TN GEMM
// D[128 x Cin] += gz^T x over all pixels (used for the GLU gate weight gradient, bnglu_tc5.cu).
template <int CIN, int TT, int TF, int NDX = 3>
__global__ void __launch_bounds__(WG_THREADS, 1)
conv_wgrad_tc5_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX,
                      float* __restrict__ gwp, int T, int F, int total_tiles) {
    pdl_enter();
    static_assert(TT * TF == WG_KPIX, "K tile must hold 32 pixels");
    constexpr int COUT = 128;
    constexpr int NCH = CIN / 32;                                   // 32-channel chunks of x
    constexpr int A_BYTES = 4 * WG_CHUNK_BYTES;                     // 128 co
    constexpr int B_BYTES = NDX * NCH * WG_CHUNK_BYTES;             // NDX taps
    constexpr int STAGE = A_BYTES + B_BYTES;
    constexpr uint32_t TMEM_COLS = NDX * CIN <= 128 ? 128 : (NDX * CIN <= 256 ? 256 : 512);
    constexpr int SHIFT = NDX == 3 ? 1 : 0;                         // tap offset = index - SHIFT
    // D fp32, A/B tf32, A and B MN-major (bits 15, 16), N = CIN, M = 128
    constexpr uint32_t IDESC = tc_idesc(CIN) | (1u << 15) | (1u << 16);
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* aligned = smem_raw + (base - smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(aligned + (size_t)WG_STAGES * STAGE);
    uint64_t* full = bars;
    uint64_t* empty = bars + WG_STAGES;
    uint64_t* accum = bars + 2 * WG_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WG_STAGES + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dy = blockIdx.y;                                      // kernel row handled by this CTA
    const int nTf = (F + TF - 1) / TF, nTt = (T + TT - 1) / TT;
    if (tid == 0) {
        for (int s = 0; s < WG_STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(accum, 1);
        fence_mbar_init();
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmG) : "memory");
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmX) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
                const int s = it % WG_STAGES, ph = (it / WG_STAGES) & 1;
                int r = tile;
                const int b = r / (nTt * nTf);
                r -= b * nTt * nTf;
                const int t0 = (r / nTf) * TT, f0 = (r % nTf) * TF;
                mbar_wait_u32(smem_u32(&empty[s]), ph ^ 1);
                const uint32_t a_dst = base + s * STAGE, b_dst = a_dst + A_BYTES;
                mbar_expect_tx(&full[s], STAGE);
#pragma unroll
                for (int c = 0; c < 4; c++)
                    tma_load_4d(a_dst + c * WG_CHUNK_BYTES, &tmG, smem_u32(&full[s]), c * 32, f0, t0, b);
#pragma unroll
                for (int dx = 0; dx < NDX; dx++)
#pragma unroll
                    for (int c = 0; c < NCH; c++)
                        tma_load_4d(b_dst + (dx * NCH + c) * WG_CHUNK_BYTES, &tmX, smem_u32(&full[s]), c * 32, f0 + dx - SHIFT,
                                    t0 + dy - SHIFT, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int it = 0; it < my_tiles; it++) {
                const int s = it % WG_STAGES, ph = (it / WG_STAGES) & 1;
                mbar_wait_u32(smem_u32(&full[s]), ph);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint32_t a_src = base + s * STAGE, b_src = a_src + A_BYTES;
#pragma unroll
                for (int k = 0; k < WG_KPIX / 8; k++) {             // 8 pixels = one 1024-byte swizzle atom per MMA
                    const uint64_t da = umma_desc_mn_sw128(a_src + k * 1024, WG_CHUNK_BYTES);
#pragma unroll
                    for (int dx = 0; dx < NDX; dx++) {
                        const uint64_t db = umma_desc_mn_sw128(b_src + dx * NCH * WG_CHUNK_BYTES + k * 1024, WG_CHUNK_BYTES);
                        umma_tf32(tmem + (uint32_t)(dx * CIN), da, db, IDESC, (it | k) != 0 ? 1u : 0u);
                    }
                }
                umma_commit(smem_u32(&empty[s]));
            }
            umma_commit(smem_u32(accum));
        }
    } else if (my_tiles > 0) {
        const int q = warp & 3;
        const int co = q * 32 + lane;
        mbar_wait_u32(smem_u32(accum), 0);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll 1
        for (int dx = 0; dx < NDX; dx++) {
            float* grow = gwp + ((size_t)(dy * NDX + dx) * COUT + co) * CIN;
#pragma unroll 1
            for (int c = 0; c < NCH; c++) {
                uint32_t v[32];
                tmem_ld32(v, tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(dx * CIN + c * 32));
#pragma unroll
                for (int j = 0; j < 8; j++)
                    atomicAdd(reinterpret_cast<float4*>(grow + c * 32) + j,
                              make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                          __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])));
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}

template <int CIN, int COUT, int TT, int TF>
int run_tc5(const CUtensorMap& tmA, const CUtensorMap& tmB, const float* bias, float* out, double* stats, int B, int T,
            int F, int cmod, cudaStream_t s) {
    auto kern = conv3x3_tc5_kernel<CIN, COUT, TT, TF>;
    static bool cfg = false;
    if (!cfg) {
        int rc = opt_in_smem(kern, TC_SMEM);
        if (rc) return rc;
        cfg = true;
    }
    dim3 grid(B * cdiv(T, TT) * cdiv(F, TF));
    SEDK_CUDA(pdl_launch(kern, dim3(grid), dim3(TC_THREADS), (size_t)(TC_SMEM), s, tmA, tmB, bias, out, stats, T, F, cmod));
    SEDK_LAUNCH_CHECK("conv3x3_tc5_kernel");
    return SEDK_OK;
}

template <int CIN, int COUT>
int run_tc5_tiles(const float* in, const float* wp, const float* bias, float* out, double* stats, int B, int T, int F,
                  int cmod, cudaStream_t s) {
    EncodeTiledFn enc = encode_fn();
    SEDK_REQUIRE(enc != nullptr, "conv3x3_tc5: cuTensorMapEncodeTiled is not available from the driver");
    SEDK_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(wp) & 15) == 0,
                 "conv3x3_tc5: operands must be 16-byte aligned");
    int TT, TF;
    if (F > 8) { TT = 8; TF = 16; }
    else if (F > 4) { TT = 16; TF = 8; }
    else if (F > 2) { TT = 32; TF = 4; }
    else { TT = 64; TF = 2; }
    CUtensorMap tmA, tmB;
    {
        cuuint64_t dims[4] = {(cuuint64_t)CIN, (cuuint64_t)F, (cuuint64_t)T, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)CIN * 4, (cuuint64_t)F * CIN * 4, (cuuint64_t)T * F * CIN * 4};
        cuuint32_t box[4] = {32, (cuuint32_t)TF, (cuuint32_t)TT, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(in), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SEDK_REQUIRE(r == CUDA_SUCCESS, "conv3x3_tc5: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)CIN, (cuuint64_t)COUT, 9};
        cuuint64_t strides[2] = {(cuuint64_t)CIN * 4, (cuuint64_t)COUT * CIN * 4};
        cuuint32_t box[3] = {32, (cuuint32_t)COUT, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(wp), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SEDK_REQUIRE(r == CUDA_SUCCESS, "conv3x3_tc5: cuTensorMapEncodeTiled(W) failed with %d", (int)r);
    }
    if (TF == 16) return run_tc5<CIN, COUT, 8, 16>(tmA, tmB, bias, out, stats, B, T, F, cmod, s);
    if (TF == 8) return run_tc5<CIN, COUT, 16, 8>(tmA, tmB, bias, out, stats, B, T, F, cmod, s);
    if (TF == 4) return run_tc5<CIN, COUT, 32, 4>(tmA, tmB, bias, out, stats, B, T, F, cmod, s);
    return run_tc5<CIN, COUT, 64, 2>(tmA, tmB, bias, out, stats, B, T, F, cmod, s);
}

template <int CIN, int TT, int TF, int NDX>
int run_wgrad_tc5(const CUtensorMap& tmG, const CUtensorMap& tmX, float* gwp, int B, int T, int F, cudaStream_t s) {
    auto kern = conv_wgrad_tc5_kernel<CIN, TT, TF, NDX>;
    constexpr size_t smem = (size_t)WG_STAGES * (4 + NDX * (CIN / 32)) * WG_CHUNK_BYTES + 1024 + 256;
    static bool cfg = false;
    if (!cfg) {
        int rc = opt_in_smem(kern, smem);
        if (rc) return rc;
        cfg = true;
    }
    const int tiles = B * cdiv(T, TT) * cdiv(F, TF);
    int gx = num_sms() / NDX;
    if (gx > tiles) gx = tiles;
    dim3 grid(gx, NDX);
    SEDK_CUDA(pdl_launch(kern, dim3(grid), dim3(WG_THREADS), (size_t)(smem), s, tmG, tmX, gwp, T, F, tiles));
    SEDK_LAUNCH_CHECK("conv_wgrad_tc5_kernel");
    return SEDK_OK;
}

template <int CIN, int NDX>
int run_wgrad_tc5_tiles(const float* x, const float* gz, float* gwp, int B, int T, int F, cudaStream_t s) {
    EncodeTiledFn enc = encode_fn();
    SEDK_REQUIRE(enc != nullptr, "conv_wgrad_tc5: cuTensorMapEncodeTiled is not available from the driver");
    int TT, TF;
    if (F >= 16) { TT = 2; TF = 16; }
    else if (F > 4) { TT = 4; TF = 8; }
    else if (F > 2) { TT = 8; TF = 4; }
    else { TT = 16; TF = 2; }
    CUtensorMap tmG, tmX;
    cuuint32_t box[4] = {32, (cuuint32_t)TF, (cuuint32_t)TT, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    {
        cuuint64_t dims[4] = {128, (cuuint64_t)F, (cuuint64_t)T, (cuuint64_t)B};
        cuuint64_t strides[3] = {128 * 4, (cuuint64_t)F * 128 * 4, (cuuint64_t)T * F * 128 * 4};
        CUresult r = enc(&tmG, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(gz), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SEDK_REQUIRE(r == CUDA_SUCCESS, "conv_wgrad_tc5: cuTensorMapEncodeTiled(gz) failed with %d", (int)r);
    }
    {
        cuuint64_t dims[4] = {(cuuint64_t)CIN, (cuuint64_t)F, (cuuint64_t)T, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)CIN * 4, (cuuint64_t)F * CIN * 4, (cuuint64_t)T * F * CIN * 4};
        CUresult r = enc(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SEDK_REQUIRE(r == CUDA_SUCCESS, "conv_wgrad_tc5: cuTensorMapEncodeTiled(x) failed with %d", (int)r);
    }
    if (TF == 16) return run_wgrad_tc5<CIN, 2, 16, NDX>(tmG, tmX, gwp, B, T, F, s);
    if (TF == 8) return run_wgrad_tc5<CIN, 4, 8, NDX>(tmG, tmX, gwp, B, T, F, s);
    if (TF == 4) return run_wgrad_tc5<CIN, 8, 4, NDX>(tmG, tmX, gwp, B, T, F, s);
    return run_wgrad_tc5<CIN, 16, 2, NDX>(tmG, tmX, gwp, B, T, F, s);
}

}  // namespace

static int g_tc5_on = -1;
bool tc5_enabled() {
    if (g_tc5_on < 0) {
        const char* e = getenv("SEDK_DISABLE_TCGEN05");
        g_tc5_on = (e != nullptr && e[0] == '1') ? 0 : 1;
    }
    return g_tc5_on == 1;
}
void tc5_set(int on) { g_tc5_on = on ? 1 : 0; }

bool tc5_supports(int cin, int cout) {
    return (cin == 32 || cin == 64 || cin == 128) && (cout == 32 || cout == 64 || cout == 128);
}

int launch_conv3x3_tc5(const float* in, const float* wp, const float* bias, float* out, double* stats, int B, int T,
                       int F, int cin, int cout, int cmod, cudaStream_t s) {
    char pname[64];
    snprintf(pname, sizeof(pname), cmod == cout ? "conv3x3_tc5_%dto%d_F%d" : "conv3x3_tc5_pair_%dto%d_F%d", cin, cout, F);
    SEDK_PROF(pname, s);
#define SEDK_TC5(CI, CO) \
    if (cin == CI && cout == CO) return run_tc5_tiles<CI, CO>(in, wp, bias, out, stats, B, T, F, cmod, s);
    SEDK_TC5(128, 128) SEDK_TC5(64, 128) SEDK_TC5(128, 64) SEDK_TC5(32, 64) SEDK_TC5(64, 32) SEDK_TC5(64, 64)
    SEDK_TC5(32, 32) SEDK_TC5(128, 32) SEDK_TC5(32, 128)
#undef SEDK_TC5
    SEDK_UNSUPPORTED("conv3x3_tc5: (cin=%d, cout=%d) has no tcgen05 instantiation", cin, cout);
}

bool tc5_wgrad_supports(int cin, int cout) { return cout == 128 && (cin == 64 || cin == 128); }

int launch_conv_wgrad_tc5(const float* x, const float* gz, float* gwpack, int B, int T, int F, int cin, int cout,
                          cudaStream_t s) {
    char pname[64];
    snprintf(pname, sizeof(pname), "conv_wgrad_tc5_%dto%d_F%d", cin, cout, F);
    SEDK_PROF(pname, s);
    SEDK_REQUIRE(tc5_wgrad_supports(cin, cout), "conv_wgrad_tc5: (cin=%d, cout=%d) not supported", cin, cout);
    if (cin == 128) return run_wgrad_tc5_tiles<128, 3>(x, gz, gwpack, B, T, F, s);
    return run_wgrad_tc5_tiles<64, 3>(x, gz, gwpack, B, T, F, s);
}

// out[n][k] += sum_pixels g[pix][n] * x[pix][k]  (both [B,T,F,128] channels-last; out pre-zeroed by the caller)
int launch_tn_gemm_tc5_c128(const float* x, const float* g, float* out, int B, int T, int F, cudaStream_t s) {
    SEDK_PROF("glu_wgrad_tc5_c128", s);
    return run_wgrad_tc5_tiles<128, 1>(x, g, out, B, T, F, s);
}

}  // namespace sedk

extern "C" int sedk_set_tcgen05(int on) {
    sedk::tc5_set(on);
    return SEDK_OK;
}
extern "C" int sedk_get_tcgen05(void) { return sedk::tc5_enabled() ? 1 : 0; }
