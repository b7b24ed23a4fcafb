// BatchNorm -> GLU -> Dropout -> AvgPool(1,2) for the 128-channel layers on tcgen05 / TMEM / TMA (sm_100a), TF32 mode.
// Reference: desed_task/nnet/CNN.py:5-16 (GLU = Linear_{C->C}(y) * sigmoid(y) over the channel axis), :73-98.
//
// The 1x1 gate GEMM is computed TRANSPOSED, channels on the MMA M axis and pixels on N:
//     lin^T[n, px] = sum_k W'[n, k] z[px, k],   W'[n, k] = Wg[n, k] * scale[k],  b'[n] = bg[n] + sum_k Wg[n, k] * shift[k]
// (BatchNorm's affine map is folded into the weight by a tiny prep kernel, so the conv output z is consumed RAW, straight from
// a 4-D TMA tensor map of the channels-last activation: box {32 ch, TF, TT, 1}, 128-byte swizzle, K-major).  With channels on
// the TMEM lanes an epilogue thread owns ONE channel and walks over pixels: every global access of a warp is a contiguous
// 128-byte row segment (z / lin / out are channels-last), the per-channel constants live in registers, the AvgPool(1,2)
// partner is the thread's next column, and per-channel reductions are plain register sums.
//
// Warp roles (576 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected lane), warps 2-17 =
// epilogue (TMEM lane quadrant = warp % 4, the four warps of a quadrant take 32 of the 128 pixel columns of a tile each).
// A tile is TT full rows of the [B, T, F, C] map (TF = F), so its 128 pixels are CONTIGUOUS in memory: all epilogue
// addressing is base + column * C.
// Persistent CTAs; z tiles double-buffered in shared memory (2 x 64 KB) next to the resident W' (64 KB); accumulators
// double-buffered in TMEM (2 x 128 columns) so that the MMA of tile i+1 overlaps the epilogue of tile i.
#include "kernels.h"
#include "tc5.cuh"

namespace sedk {
namespace {

constexpr int GT_C = 128;                        // channels (MMA M and K)
constexpr int GT_NPX = 128;                      // pixels per tile (MMA N)
constexpr int GT_CHUNK = GT_NPX * 128;           // bytes of one 32-channel chunk of a tile: 128 rows x 128 B
constexpr int GT_TILE = 4 * GT_CHUNK;            // 64 KB
constexpr int GT_EPI_WARPS = 16;                // 4 per TMEM lane quadrant, 32 pixel columns each
constexpr int GT_THREADS = 64 + 32 * GT_EPI_WARPS;
constexpr size_t GT_SMEM_FWD = (size_t)3 * GT_TILE + 1024 + 256;

// ------------------------------------------------------------------------------------------------------------------------
// bn_finalize + gate-weight preparation: one block per gate output n, one thread per input channel k.
//   pack[0 .. C*C)      W'[n][k] = Wg[n][k] * scale[k]        (forward A operand)
//   pack[C*C .. 2*C*C)  WT[k][n] = Wg[n][k]                   (backward A operand: g_y^T = WT g_lin^T)
//   pack[2*C*C .. +C)   b'[n]    = bg[n] + sum_k Wg[n][k] * shift[k]
__global__ void __launch_bounds__(GT_C)
glu_prep_kernel(const double* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                float* __restrict__ rm, float* __restrict__ rv, int64_t* __restrict__ nb, float* __restrict__ bn,
                const float* __restrict__ glu_w, const float* __restrict__ glu_b, float* __restrict__ pack, double count,
                float eps, float momentum, int training) {
    constexpr int C = GT_C;
    __shared__ float red[C / 32];
    const int n = blockIdx.x, k = threadIdx.x;
    float mean, invstd;
    double unbiased = 0.0;
    if (training) {
        const double m = stats[k] / count;
        double var = stats[C + k] / count - m * m;
        if (var < 0.0) var = 0.0;
        mean = (float)m;
        invstd = (float)(1.0 / sqrt(var + (double)eps));
        unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    } else {
        mean = rm[k];
        invstd = 1.0f / sqrtf(rv[k] + eps);
    }
    const float scale = gamma[k] * invstd;
    const float shift = beta[k] - mean * scale;
    const float w = glu_w[n * C + k];
    // rounded to the nearest TF32 value here: the tensor core truncates what it reads
    pack[n * C + k] = __uint_as_float(to_tf32(w * scale));
    pack[C * C + k * C + n] = __uint_as_float(to_tf32(w));
    float part = warp_sum(w * shift);
    if ((k & 31) == 0) red[k >> 5] = part;
    __syncthreads();
    if (k == 0) {
        float s = glu_b[n];
        for (int i = 0; i < C / 32; i++) s += red[i];
        pack[2 * C * C + n] = s;
    }
    if (n == 0) {
        // the last reader of the running statistics is this block itself (all other blocks only read stats / gamma / beta)
        if (training) {
            rm[k] = (1.0f - momentum) * rm[k] + momentum * mean;
            rv[k] = (1.0f - momentum) * rv[k] + momentum * (float)unbiased;
            if (k == 0 && nb != nullptr) *nb += 1;
        }
        bn[k] = scale;
        bn[C + k] = shift;
        bn[2 * C + k] = mean;
        bn[3 * C + k] = invstd;
    }
}

struct GtGeom {
    int T, F, TT, TF, nTt, To, Fo;      // pooling (1, 2): To = T, Fo = F / 2; one tile column (TF == F)
    int tf_shift;
};

// byte offset of element (pixel row p, channel-in-chunk `lane`) inside a 128-byte-swizzled [128 rows x 32 fp32] chunk
__device__ __forceinline__ uint32_t sw128_off(int p, int lane) {
    return (uint32_t)(p * 128 + ((((lane >> 2) ^ (p & 7)) << 4) | ((lane & 3) << 2)));
}

// keep / drop of the 8 pixels [8 g8, 8 g8 + 8) of channel n in tile `tile`: 16-bit draws, bit e of the result = keep pixel e.
// Shared by the forward and the backward kernel of this file (masks are regenerated, never stored).
__device__ __forceinline__ uint32_t gt_keep8(const Philox& ph, int tile, int g8, int n, uint64_t dstream, uint32_t thresh16) {
    const uint4 r = ph(((uint64_t)tile * 16ull + (uint64_t)g8) * (uint64_t)GT_C + (uint64_t)n, dstream);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    uint32_t bits = 0;
#pragma unroll
    for (int e = 0; e < 4; e++) {
        bits |= ((w[e] & 0xffffu) >= thresh16 ? 1u : 0u) << (2 * e);
        bits |= ((w[e] >> 16) >= thresh16 ? 1u : 0u) << (2 * e + 1);
    }
    return bits;
}

__global__ void __launch_bounds__(GT_THREADS, 1)
bnglu_tc5_fwd_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmW,
                     const float* __restrict__ bn, const float* __restrict__ bprime, float* __restrict__ out,
                     float* __restrict__ lin_out, GtGeom gm, int total_tiles, uint32_t thresh16, float inv_keep,
                     uint64_t seed, const uint64_t* __restrict__ seed_dev, uint64_t dstream) {
    constexpr int C = GT_C;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* aligned = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t w_smem = base;                               // W' : 4 chunks [128 n x 32 k]
    const uint32_t z_smem = base + GT_TILE;                     // z  : 2 stages x 4 chunks [128 px x 32 k]
    uint64_t* bars = reinterpret_cast<uint64_t*>(aligned + (size_t)3 * GT_TILE);
    uint64_t* wfull = bars;            // [1]
    uint64_t* zfull = bars + 1;        // [2]
    uint64_t* zempty = bars + 3;       // [2]
    uint64_t* accfull = bars + 5;      // [2]
    uint64_t* accempty = bars + 7;     // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
    constexpr uint32_t IDESC = tc_idesc(GT_NPX);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        mbar_init(wfull, 1);
        for (int s = 0; s < 2; s++) {
            mbar_init(&zfull[s], 1);
            mbar_init(&zempty[s], GT_EPI_WARPS);
            mbar_init(&accfull[s], 1);
            mbar_init(&accempty[s], GT_EPI_WARPS);
        }
        fence_mbar_init();
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmZ) : "memory");
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmW) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                     "r"(256)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc5_fence_before();
    __syncthreads();
    tc5_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(wfull, GT_TILE);
            for (int c = 0; c < 4; c++) {
                asm volatile(
                    "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
                        w_smem + c * GT_CHUNK),
                    "l"(&tmW), "r"(smem_u32(wfull)), "r"(c * 32), "r"(0)
                    : "memory");
            }
            for (int it = 0; it < my_tiles; it++) {
                const int tile = blockIdx.x + it * gridDim.x;
                const int s = it & 1, ph = (it >> 1) & 1;
                const int b = tile / gm.nTt, t0 = (tile - b * gm.nTt) * gm.TT;
                mbar_wait_u32(smem_u32(&zempty[s]), ph ^ 1);
                mbar_expect_tx(&zfull[s], GT_TILE);
                for (int c = 0; c < 4; c++)
                    tma_load_4d(z_smem + s * GT_TILE + c * GT_CHUNK, &tmZ, smem_u32(&zfull[s]), c * 32, 0, t0, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            mbar_wait_u32(smem_u32(wfull), 0);
            for (int it = 0; it < my_tiles; it++) {
                const int s = it & 1, ph = (it >> 1) & 1;
                mbar_wait_u32(smem_u32(&zfull[s]), ph);
                mbar_wait_u32(smem_u32(&accempty[s]), ph ^ 1);
                tc5_fence_after();
                const uint32_t d = tmem + (uint32_t)(s * GT_NPX);
#pragma unroll
                for (int c = 0; c < 4; c++)
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        umma_tf32(d, umma_desc_sw128(w_smem + c * GT_CHUNK + k * 32),
                                  umma_desc_sw128(z_smem + s * GT_TILE + c * GT_CHUNK + k * 32), IDESC, (c | k) != 0 ? 1u : 0u);
                umma_commit(smem_u32(&accfull[s]));
            }
        }
    } else {
        // ---- epilogue: this thread owns channel n = 32 q + lane (TMEM lane) and the pixel columns [c0, c0 + 32)
        const int q = warp & 3, c0 = 32 * ((warp - 2) >> 2);
        const int n = 32 * q + lane;
        const float sc = bn[n], sh = bn[C + n], bp = bprime[n];
        const Philox ph(seed + (seed_dev ? *seed_dev : 0ull));
        const uint32_t swz_hi = (uint32_t)(lane >> 2), swz_lo = (uint32_t)((lane & 3) << 2);
        for (int it = 0; it < my_tiles; it++) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int s = it & 1, phs = (it >> 1) & 1;
            const int b = tile / gm.nTt, t0 = (tile - b * gm.nTt) * gm.TT;
            const int pvalid = (gm.T - t0) * gm.F - c0;               // columns i < pvalid of this warp's 32 are real pixels
            const size_t row0 = (size_t)b * gm.T + t0;
            float* lrow = lin_out != nullptr ? lin_out + (row0 * gm.F + c0) * C + n : nullptr;
            float* orow = out + (row0 * gm.Fo + (c0 >> 1)) * C + n;
            mbar_wait_u32(smem_u32(&accfull[s]), phs);
            tc5_fence_after();
            const uint8_t* zs = aligned + GT_TILE + (size_t)s * GT_TILE + (size_t)q * GT_CHUNK + (size_t)c0 * 128;
            uint32_t v[32];
            tmem_ld32(v, tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * GT_NPX + c0));
#pragma unroll
            for (int g8 = 0; g8 < 4; g8++) {
                uint32_t kb = 0xffu;
                if (thresh16 != 0u) kb = gt_keep8(ph, tile, (c0 >> 3) + g8, n, dstream, thresh16);
#pragma unroll
                for (int e = 0; e < 8; e += 2) {
                    const int i = 8 * g8 + e;                           // even pixel of a pooling pair
                    float r[2];
#pragma unroll
                    for (int d = 0; d < 2; d++) {
                        const float z = *reinterpret_cast<const float*>(
                            zs + (i + d) * 128 + (((swz_hi ^ (uint32_t)((i + d) & 7)) << 4) | swz_lo));
                        const float lin = __uint_as_float(v[i + d]) + bp;
                        float a = lin * fast_sigmoidf_(fmaf(z, sc, sh));
                        if (thresh16 != 0u) a = ((kb >> (e + d)) & 1u) ? a * inv_keep : 0.f;
                        r[d] = a;
                        if (lrow != nullptr && i + d < pvalid) lrow[(size_t)(i + d) * C] = lin;
                    }
                    // the pooled activation is the next convolution's MMA operand: store the nearest TF32 value
                    if (i < pvalid) orow[(size_t)(i >> 1) * C] = __uint_as_float(to_tf32(0.5f * (r[0] + r[1])));
                }
            }
            tc5_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(smem_u32(&accempty[s]));
                mbar_arrive(smem_u32(&zempty[s]));
            }
        }
    }
    tc5_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc5_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(256) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// Backward.  The forward saved lin = W' z + b' (gate pre-activation), so only the data GEMM is left in this kernel:
//     g_y^T[k, px] = e^T[k, px] + sum_n WT[k, n] g_lin[px, n]        (WT = Wg^T, K-major A operand, resident in smem)
// Per 128-pixel tile (single-buffered; three phases):
//   A  epilogue threads (thread = channel n): lin / gout from global (coalesced), z from the TMA-staged tile -> sigmoid,
//      dropout mask, g_lin, elementwise term e.  g_lin goes to shared memory in the 128-byte-swizzled K-major layout the MMA
//      reads as B AND to global memory (over lin) for the weight-gradient GEMM; e is written into the TMEM accumulator with
//      tcgen05.st, so the MMA simply accumulates on top of it.
//   B  MMA warp: 16 x tcgen05.mma (M = 128 channels, N = 128 pixels, K = 8), commit.
//   C  epilogue threads: tcgen05.ld g_y^T -> coalesced store of g_y, per-channel sum g_y and sum g_y * zhat in registers.
// The gate weight gradient dWg = g_lin^T y is a separate TN GEMM over all pixels on the side stream
// (launch_tn_gemm_tc5_c128: tcgen05, MN-major operands straight from global g_lin and RAW z) followed by glu_wgrad_fix_kernel:
//     dWg[n][k] = scale[k] * (g_lin^T z)[n][k] + shift[k] * sum_px g_lin[px][n].
constexpr size_t GT_SMEM_BWD = (size_t)3 * GT_TILE + 1024 + 256;

__global__ void __launch_bounds__(GT_THREADS, 1)
bnglu_tc5_bwd_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmW,
                     const float* __restrict__ bn, const float* __restrict__ gout, float* __restrict__ lin_glin,
                     float* __restrict__ gy, float* __restrict__ gglu_b, double* __restrict__ stats, GtGeom gm,
                     int total_tiles, uint32_t thresh16, float inv_keep, uint64_t seed,
                     const uint64_t* __restrict__ seed_dev, uint64_t dstream) {
    constexpr int C = GT_C;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* aligned = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t w_smem = base;                               // WT : 4 chunks [128 k x 32 n]
    const uint32_t z_smem = base + GT_TILE;                     // z  : 4 chunks [128 px x 32 k]
    const uint32_t g_smem = base + 2 * GT_TILE;                 // g_lin : 4 chunks [128 px x 32 n]
    uint64_t* bars = reinterpret_cast<uint64_t*>(aligned + (size_t)3 * GT_TILE);
    uint64_t* wfull = bars;
    uint64_t* zfull = bars + 1;
    uint64_t* zempty = bars + 2;
    uint64_t* gfull = bars + 3;
    uint64_t* dfull = bars + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
    constexpr uint32_t IDESC = tc_idesc(GT_NPX);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        mbar_init(wfull, 1);
        mbar_init(zfull, 1);
        mbar_init(zempty, GT_EPI_WARPS);
        mbar_init(gfull, GT_EPI_WARPS);
        mbar_init(dfull, 1);
        fence_mbar_init();
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmZ) : "memory");
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmW) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                     "r"(128)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc5_fence_before();
    __syncthreads();
    tc5_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(wfull, GT_TILE);
            for (int c = 0; c < 4; c++) {
                asm volatile(
                    "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
                        w_smem + c * GT_CHUNK),
                    "l"(&tmW), "r"(smem_u32(wfull)), "r"(c * 32), "r"(0)
                    : "memory");
            }
            for (int it = 0; it < my_tiles; it++) {
                const int tile = blockIdx.x + it * gridDim.x;
                const int b = tile / gm.nTt, t0 = (tile - b * gm.nTt) * gm.TT;
                mbar_wait_u32(smem_u32(zempty), (it & 1) ^ 1);
                mbar_expect_tx(zfull, GT_TILE);
                for (int c = 0; c < 4; c++) tma_load_4d(z_smem + c * GT_CHUNK, &tmZ, smem_u32(zfull), c * 32, 0, t0, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            mbar_wait_u32(smem_u32(wfull), 0);
            for (int it = 0; it < my_tiles; it++) {
                mbar_wait_u32(smem_u32(gfull), it & 1);
                tc5_fence_after();
#pragma unroll
                for (int c = 0; c < 4; c++)
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        umma_tf32(tmem, umma_desc_sw128(w_smem + c * GT_CHUNK + k * 32),
                                  umma_desc_sw128(g_smem + c * GT_CHUNK + k * 32), IDESC, 1u);
                umma_commit(smem_u32(dfull));
            }
        }
    } else {
        const int q = warp & 3, c0 = 32 * ((warp - 2) >> 2);
        const int n = 32 * q + lane;
        const float sc = bn[n], sh = bn[C + n], is = bn[3 * C + n], mi = -bn[2 * C + n] * bn[3 * C + n];
        const Philox ph(seed + (seed_dev ? *seed_dev : 0ull));
        const uint32_t swz_hi = (uint32_t)(lane >> 2), swz_lo = (uint32_t)((lane & 3) << 2);
        float s_gy = 0.f, s_gyz = 0.f, s_gl = 0.f;
        const uint8_t* zs = aligned + GT_TILE + (size_t)q * GT_CHUNK + (size_t)c0 * 128;
        uint8_t* gs = aligned + 2 * GT_TILE + (size_t)q * GT_CHUNK + (size_t)c0 * 128;
        for (int it = 0; it < my_tiles; it++) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int b = tile / gm.nTt, t0 = (tile - b * gm.nTt) * gm.TT;
            const int pvalid = (gm.T - t0) * gm.F - c0;
            const size_t row0 = (size_t)b * gm.T + t0;
            float* lrow = lin_glin + (row0 * gm.F + c0) * C + n;
            const float* grow = gout + (row0 * gm.Fo + (c0 >> 1)) * C + n;
            float* yrow = gy + (row0 * gm.F + c0) * C + n;
            // ---------------- phase A (global operands are fetched in batches of 16 pixels before they are consumed)
            uint32_t ev[32];
#pragma unroll
            for (int hb = 0; hb < 2; hb++) {
                float linv[16], gov[8];
#pragma unroll
                for (int i = 0; i < 16; i++) linv[i] = (16 * hb + i < pvalid) ? lrow[(size_t)(16 * hb + i) * C] : 0.f;
#pragma unroll
                for (int i = 0; i < 8; i++) gov[i] = (16 * hb + 2 * i < pvalid) ? grow[(size_t)(8 * hb + i) * C] : 0.f;
                if (hb == 0) mbar_wait_u32(smem_u32(zfull), it & 1);
#pragma unroll
                for (int g8 = 0; g8 < 2; g8++) {
                    uint32_t kb = 0xffu;
                    if (thresh16 != 0u) kb = gt_keep8(ph, tile, (c0 >> 3) + 2 * hb + g8, n, dstream, thresh16);
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        const int i = 16 * hb + 8 * g8 + e;
                        const uint32_t so = (uint32_t)(i * 128) + (((swz_hi ^ (uint32_t)(i & 7)) << 4) | swz_lo);
                        float ga = 0.5f * gov[(8 * g8 + e) >> 1];
                        if (thresh16 != 0u) ga = ((kb >> e) & 1u) ? ga * inv_keep : 0.f;
                        const float z = *reinterpret_cast<const float*>(zs + so);
                        const float sg = fast_sigmoidf_(fmaf(z, sc, sh));
                        const float g_lin = ga * sg;                       // 0 for padding pixels (gov = 0)
                        ev[i] = __float_as_uint(ga * linv[8 * g8 + e] * sg * (1.0f - sg));
                        *reinterpret_cast<float*>(gs + so) = g_lin;
                        if (i < pvalid) lrow[(size_t)i * C] = g_lin;
                        s_gl += g_lin;
                    }
                }
            }
            tmem_st32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, ev);
            fence_proxy_async();          // the generic-proxy writes of g_lin must be visible to the tensor core (async proxy)
            tc5_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(gfull));
            // ---------------- phase C
            mbar_wait_u32(smem_u32(dfull), it & 1);
            tc5_fence_after();
            uint32_t v[32];
            tmem_ld32(v, tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0);
#pragma unroll
            for (int i = 0; i < 32; i++) {
                if (i < pvalid) {
                    const float g = __uint_as_float(v[i]);
                    yrow[(size_t)i * C] = g;
                    const float z = *reinterpret_cast<const float*>(
                        zs + (uint32_t)(i * 128) + (((swz_hi ^ (uint32_t)(i & 7)) << 4) | swz_lo));
                    s_gy += g;
                    s_gyz = fmaf(g, fmaf(z, is, mi), s_gyz);
                }
            }
            tc5_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(zempty));
        }
        atomicAdd(&stats[2 * C + n], (double)s_gy);
        atomicAdd(&stats[3 * C + n], (double)s_gyz);
        atomicAdd(&gglu_b[n], s_gl);
    }
    tc5_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc5_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(128) : "memory");
    }
}

// dWg[n][k] = scale[k] * raw[n][k] + shift[k] * sum_px g_lin[px][n]   (in place; raw = g_lin^T z, gglu_b = sum g_lin)
__global__ void glu_wgrad_fix_kernel(float* __restrict__ gglu_w, const float* __restrict__ gglu_b, const float* __restrict__ bn) {
    constexpr int C = GT_C;
    const int n = blockIdx.x, k = threadIdx.x;
    gglu_w[n * C + k] = fmaf(bn[k], gglu_w[n * C + k], bn[C + k] * gglu_b[n]);
}

inline bool make_gtgeom(GtGeom& g, int T, int F, int pt, int pf) {
    if (pt != 1 || pf != 2) return false;
    if (F != 16 && F != 8 && F != 4 && F != 2) return false;
    g.T = T; g.F = F; g.TF = F; g.TT = GT_NPX / F;
    g.nTt = cdiv(T, g.TT);
    g.To = T; g.Fo = F / 2;
    g.tf_shift = 0;
    while ((1 << g.tf_shift) < F) g.tf_shift++;
    return true;
}

inline uint32_t drop_threshold16(float p) {
    if (p <= 0.f) return 0u;
    double t = (double)p * 65536.0 + 0.5;
    if (t < 1.0) t = 1.0;
    if (t > 65535.0) t = 65535.0;
    return (uint32_t)t;
}

int make_maps(const float* z, const float* wmat, int B, const GtGeom& gm, CUtensorMap* tmZ, CUtensorMap* tmW) {
    EncodeTiledFn enc = encode_fn();
    SEDK_REQUIRE(enc != nullptr, "bnglu_tc5: cuTensorMapEncodeTiled is not available from the driver");
    SEDK_REQUIRE((reinterpret_cast<uintptr_t>(z) & 15) == 0 && (reinterpret_cast<uintptr_t>(wmat) & 15) == 0,
                 "bnglu_tc5: operands must be 16-byte aligned");
    {
        cuuint64_t dims[4] = {(cuuint64_t)GT_C, (cuuint64_t)gm.F, (cuuint64_t)gm.T, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)GT_C * 4, (cuuint64_t)gm.F * GT_C * 4, (cuuint64_t)gm.T * gm.F * GT_C * 4};
        cuuint32_t box[4] = {32, (cuuint32_t)gm.TF, (cuuint32_t)gm.TT, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(tmZ, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(z), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SEDK_REQUIRE(r == CUDA_SUCCESS, "bnglu_tc5: cuTensorMapEncodeTiled(z) failed with %d", (int)r);
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)GT_C, (cuuint64_t)GT_C};
        cuuint64_t strides[1] = {(cuuint64_t)GT_C * 4};
        cuuint32_t box[2] = {32, (cuuint32_t)GT_C};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(wmat), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SEDK_REQUIRE(r == CUDA_SUCCESS, "bnglu_tc5: cuTensorMapEncodeTiled(W) failed with %d", (int)r);
    }
    return SEDK_OK;
}

}  // namespace

bool bnglu_tc5_supports(int T, int F, int C, int pt, int pf, int precision) {
    GtGeom g;
    return C == GT_C && precision == 0 && tc5_enabled() && get_option("bnglu_tc5", 1) != 0 && make_gtgeom(g, T, F, pt, pf);
}

int launch_glu_prep(const double* stats, const float* gamma, const float* beta, float* running_mean, float* running_var,
                    int64_t* num_batches, float* bn, const float* glu_w, const float* glu_b, float* pack, double count,
                    float eps, float momentum, int training, int C, cudaStream_t s) {
    SEDK_PROF("glu_prep", s);
    SEDK_REQUIRE(C == GT_C, "glu_prep: C must be %d", GT_C);
    glu_prep_kernel<<<GT_C, GT_C, 0, s>>>(stats, gamma, beta, running_mean, running_var, num_batches, bn, glu_w, glu_b, pack,
                                          count, eps, momentum, training);
    SEDK_LAUNCH_CHECK("glu_prep_kernel");
    return SEDK_OK;
}

int launch_bnglu_tc5_fwd(const float* z, const float* bn, const float* pack, float* out, float* lin_out, int B, int T, int F,
                         int pt, int pf, float drop_p, uint64_t seed, const uint64_t* seed_dev, uint64_t drop_stream,
                         cudaStream_t s) {
    SEDK_PROF("bnglu_tc5_fwd_c128", s);
    GtGeom gm;
    SEDK_REQUIRE(make_gtgeom(gm, T, F, pt, pf), "bnglu_tc5: unsupported geometry");
    CUtensorMap tmZ, tmW;
    int rc = make_maps(z, pack, B, gm, &tmZ, &tmW);
    if (rc) return rc;
    static bool cfg = false;
    if (!cfg) {
        rc = opt_in_smem(bnglu_tc5_fwd_kernel, GT_SMEM_FWD);
        if (rc) return rc;
        cfg = true;
    }
    const int tiles = B * gm.nTt;
    const int grid = tiles < num_sms() ? tiles : num_sms();
    bnglu_tc5_fwd_kernel<<<grid, GT_THREADS, GT_SMEM_FWD, s>>>(tmZ, tmW, bn, pack + 2 * GT_C * GT_C, out, lin_out, gm, tiles,
                                                               drop_threshold16(drop_p),
                                                               drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f, seed, seed_dev,
                                                               drop_stream);
    SEDK_LAUNCH_CHECK("bnglu_tc5_fwd_kernel");
    return SEDK_OK;
}

int launch_bnglu_tc5_bwd(const float* z, const float* bn, const float* pack, const float* gout, float* lin_glin, float* gy,
                         float* gglu_b, double* stats, int B, int T, int F, int pt, int pf, float drop_p, uint64_t seed,
                         const uint64_t* seed_dev, uint64_t drop_stream, cudaStream_t s) {
    SEDK_PROF("bnglu_tc5_bwd_c128", s);
    GtGeom gm;
    SEDK_REQUIRE(make_gtgeom(gm, T, F, pt, pf), "bnglu_tc5: unsupported geometry");
    CUtensorMap tmZ, tmW;
    int rc = make_maps(z, pack + GT_C * GT_C, B, gm, &tmZ, &tmW);          // A operand = WT
    if (rc) return rc;
    static bool cfg = false;
    if (!cfg) {
        rc = opt_in_smem(bnglu_tc5_bwd_kernel, GT_SMEM_BWD);
        if (rc) return rc;
        cfg = true;
    }
    const int tiles = B * gm.nTt;
    const int grid = tiles < num_sms() ? tiles : num_sms();
    bnglu_tc5_bwd_kernel<<<grid, GT_THREADS, GT_SMEM_BWD, s>>>(tmZ, tmW, bn, gout, lin_glin, gy, gglu_b, stats, gm, tiles,
                                                               drop_threshold16(drop_p),
                                                               drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f, seed, seed_dev,
                                                               drop_stream);
    SEDK_LAUNCH_CHECK("bnglu_tc5_bwd_kernel");
    return SEDK_OK;
}

// gglu_w (zeroed by the caller) <- g_lin^T z on tcgen05, then the BatchNorm-fold fix-up; needs gglu_b = sum g_lin complete
int launch_glu_wgrad_tc5(const float* z, const float* g_lin, const float* bn, float* gglu_w, const float* gglu_b, int B, int T,
                         int F, cudaStream_t s) {
    int rc = launch_tn_gemm_tc5_c128(z, g_lin, gglu_w, B, T, F, s);
    if (rc) return rc;
    SEDK_PROF("glu_wgrad_fix", s);
    glu_wgrad_fix_kernel<<<GT_C, GT_C, 0, s>>>(gglu_w, gglu_b, bn);
    SEDK_LAUNCH_CHECK("glu_wgrad_fix_kernel");
    return SEDK_OK;
}

}  // namespace sedk
