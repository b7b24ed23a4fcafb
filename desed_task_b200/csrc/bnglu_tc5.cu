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
//
// 64-channel layers run on the SAME kernels: two row-halves of a tile are stacked on the M / K axes, i.e. MMA row
// m = (h, n) is channel n of the pixels in half h, the weight operand is the 128 x 128 block-diagonal diag(W, W), and the
// B operand row of pixel column px is [z(half 0, px, 0..63) | z(half 1, px, 0..63)] (two TMA boxes per K chunk pair).
// Half the MMA work multiplies zeros - irrelevant here, the layer is bound by its element-wise work and HBM traffic.
#include "kernels.h"
#include "tc5.cuh"

namespace sedk {
namespace {

constexpr int GT_C = 128;                        // MMA M and K: 128 channels, or 2 halves x 64 channels
constexpr int GT_NPX = 128;                      // pixels per tile (MMA N)
constexpr int GT_CHUNK = GT_NPX * 128;           // bytes of one 32-channel chunk of a tile: 128 rows x 128 B
constexpr int GT_TILE = 4 * GT_CHUNK;            // 64 KB
constexpr int GT_EPI_WARPS = 16;                // 4 per TMEM lane quadrant, 32 pixel columns each
constexpr int GT_THREADS = 64 + 32 * GT_EPI_WARPS;
constexpr size_t GT_SMEM_FWD = (size_t)3 * GT_TILE + 1024 + 256;

// ------------------------------------------------------------------------------------------------------------------------
// bn_finalize + gate-weight preparation: one block per gate output n, one thread per input channel k.
//   pack[0 .. 128*128)            W'[n][k] = Wg[n][k] * scale[k]       (forward A operand; block-diagonal for C = 64)
//   pack[128*128 .. 2*128*128)    WT[k][n] = Wg[n][k]                  (backward A operand: g_y^T = WT g_lin^T)
//   pack[2*128*128 .. +C)         b'[n]    = bg[n] + sum_k Wg[n][k] * shift[k]
//   pack[2*128*128+128 .. +128*128)  scratch of the gate weight-gradient GEMM (C = 64 only)
constexpr int GT_PACK_B = 2 * GT_C * GT_C;       // offset of b'
constexpr int GT_PACK_RAW = 2 * GT_C * GT_C + GT_C;
template <int C>
__global__ void __launch_bounds__(GT_C)
glu_prep_kernel(const double* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                float* __restrict__ rm, float* __restrict__ rv, int64_t* __restrict__ nb, float* __restrict__ bn,
                const float* __restrict__ glu_w, const float* __restrict__ glu_b, float* __restrict__ pack, double count,
                float eps, float momentum, int training) {
    pdl_enter();
    constexpr int M = GT_C;
    __shared__ float red[M / 32];
    // pack row m = (h, n), column j = (h', k); the blocks h != h' of the two 128 x 128 operands are zero
    const int m = blockIdx.x, j = threadIdx.x;
    const int n = m & (C - 1), k = j & (C - 1);
    const bool diag = (m / C) == (j / C);
    float mean, invstd;
    double unbiased = 0.0;
    if (training) {
        const double mu = stats[k] / count;
        double var = stats[C + k] / count - mu * mu;
        if (var < 0.0) var = 0.0;
        mean = (float)mu;
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
    pack[m * M + j] = diag ? __uint_as_float(to_tf32(w * scale)) : 0.f;
    // WT2[(h, kk)][(h', nn)] = Wg[nn][kk]: with (m, j) = ((h, kk), (h', nn)) that is glu_w[k * C + n]
    pack[M * M + m * M + j] = diag ? __uint_as_float(to_tf32(glu_w[k * C + n])) : 0.f;
    float part = warp_sum(j < C ? w * shift : 0.f);
    if ((j & 31) == 0) red[j >> 5] = part;
    __syncthreads();
    if (j == 0 && m < C) {
        float s = glu_b[n];
        for (int i = 0; i < M / 32; i++) s += red[i];
        pack[2 * M * M + n] = s;
    }
    if (m == 0 && j < C) {
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
    int T, F, TTh, TT, nTt, Fo;         // pooling (1, 2): Fo = F / 2; tile = TT full rows = (128 / C) halves of TTh rows,
                                        // TTh * F = 128 pixel columns per half
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

template <int C>
__global__ void __launch_bounds__(GT_THREADS, 1)
bnglu_tc5_fwd_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmW,
                     const float* __restrict__ bn, const float* __restrict__ bprime, float* __restrict__ out,
                     float* __restrict__ lin_out, GtGeom gm, int total_tiles, uint32_t thresh16, float inv_keep,
                     uint64_t seed, const uint64_t* __restrict__ seed_dev, uint64_t dstream) {
    pdl_enter();
    constexpr int CH = C / 32;                                  // 32-channel chunks per half
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
                for (int c = 0; c < 4; c++)     // chunk c = (half c / CH, channels 32 (c % CH) ..)
                    tma_load_4d(z_smem + s * GT_TILE + c * GT_CHUNK, &tmZ, smem_u32(&zfull[s]), (c % CH) * 32, 0,
                                t0 + (c / CH) * gm.TTh, b);
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
        // ---- epilogue: this thread owns TMEM lane m = 32 q + lane = channel n of row-half h, pixel columns [c0, c0 + 32)
        const int q = warp & 3, c0 = 32 * ((warp - 2) >> 2);
        const int m = 32 * q + lane, n = m & (C - 1), h = m / C;
        const float sc = bn[n], sh = bn[C + n], bp = bprime[n];
        const Philox ph(seed + (seed_dev ? *seed_dev : 0ull));
        const uint32_t swz_hi = (uint32_t)(lane >> 2), swz_lo = (uint32_t)((lane & 3) << 2);
        for (int it = 0; it < my_tiles; it++) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int s = it & 1, phs = (it >> 1) & 1;
            const int b = tile / gm.nTt, t0 = (tile - b * gm.nTt) * gm.TT + h * gm.TTh;
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
                if (thresh16 != 0u) kb = gt_keep8(ph, tile, (c0 >> 3) + g8, m, dstream, thresh16);
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

template <int C>
__global__ void __launch_bounds__(GT_THREADS, 1)
bnglu_tc5_bwd_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmW,
                     const float* __restrict__ bn, const float* __restrict__ gout, float* __restrict__ lin_glin,
                     float* __restrict__ gy, float* __restrict__ gglu_b, double* __restrict__ stats, GtGeom gm,
                     int total_tiles, uint32_t thresh16, float inv_keep, uint64_t seed,
                     const uint64_t* __restrict__ seed_dev, uint64_t dstream) {
    pdl_enter();
    constexpr int CH = C / 32;
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
                for (int c = 0; c < 4; c++)
                    tma_load_4d(z_smem + c * GT_CHUNK, &tmZ, smem_u32(zfull), (c % CH) * 32, 0, t0 + (c / CH) * gm.TTh, b);
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
        const int m = 32 * q + lane, n = m & (C - 1), h = m / C;
        const float sc = bn[n], sh = bn[C + n], is = bn[3 * C + n], mi = -bn[2 * C + n] * bn[3 * C + n];
        const Philox ph(seed + (seed_dev ? *seed_dev : 0ull));
        const uint32_t swz_hi = (uint32_t)(lane >> 2), swz_lo = (uint32_t)((lane & 3) << 2);
        float s_gy = 0.f, s_gyz = 0.f, s_gl = 0.f;
        const uint8_t* zs = aligned + GT_TILE + (size_t)q * GT_CHUNK + (size_t)c0 * 128;
        uint8_t* gs = aligned + 2 * GT_TILE + (size_t)q * GT_CHUNK + (size_t)c0 * 128;
        for (int it = 0; it < my_tiles; it++) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int b = tile / gm.nTt, t0 = (tile - b * gm.nTt) * gm.TT + h * gm.TTh;
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
                    if (thresh16 != 0u) kb = gt_keep8(ph, tile, (c0 >> 3) + 2 * hb + g8, m, dstream, thresh16);
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

// dWg[n][k] = scale[k] * raw[n][k] + shift[k] * sum_px g_lin[px][n]   (raw = g_lin^T z, gglu_b = sum g_lin).
// C = 128: raw IS gglu_w (in place).  C = 64: raw is the 128 x 128 product of the pixel-pair views; its two diagonal
// 64 x 64 blocks are the sums over even / odd pixels, the off-diagonal blocks are cross terms and are dropped.
template <int C>
__global__ void glu_wgrad_fix_kernel(float* __restrict__ gglu_w, const float* __restrict__ raw, const float* __restrict__ gglu_b,
                                     const float* __restrict__ bn) {
    const int n = blockIdx.x, k = threadIdx.x;
    float r;
    if (C == GT_C) r = raw[n * GT_C + k];
    else r = raw[n * GT_C + k] + raw[(C + n) * GT_C + C + k];
    gglu_w[n * C + k] = fmaf(bn[k], r, bn[C + k] * gglu_b[n]);
}

inline bool make_gtgeom(GtGeom& g, int T, int F, int C, int pt, int pf) {
    if (pt != 1 || pf != 2 || (C != 64 && C != 128)) return false;
    if (F < 2 || F > 64 || (F & (F - 1)) != 0) return false;
    g.T = T; g.F = F;
    g.TTh = GT_NPX / F;
    g.TT = g.TTh * (GT_C / C);
    g.nTt = cdiv(T, g.TT);
    g.Fo = F / 2;
    return true;
}

inline uint32_t drop_threshold16(float p) {
    if (p <= 0.f) return 0u;
    double t = (double)p * 65536.0 + 0.5;
    if (t < 1.0) t = 1.0;
    if (t > 65535.0) t = 65535.0;
    return (uint32_t)t;
}

int make_maps(const float* z, const float* wmat, int B, int C, const GtGeom& gm, CUtensorMap* tmZ, CUtensorMap* tmW) {
    EncodeTiledFn enc = encode_fn();
    SEDK_REQUIRE(enc != nullptr, "bnglu_tc5: cuTensorMapEncodeTiled is not available from the driver");
    SEDK_REQUIRE((reinterpret_cast<uintptr_t>(z) & 15) == 0 && (reinterpret_cast<uintptr_t>(wmat) & 15) == 0,
                 "bnglu_tc5: operands must be 16-byte aligned");
    {
        // one box = one 32-channel chunk of one row-half: TTh full rows x F columns = 128 pixels
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)gm.F, (cuuint64_t)gm.T, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)gm.F * C * 4, (cuuint64_t)gm.T * gm.F * C * 4};
        cuuint32_t box[4] = {32, (cuuint32_t)gm.F, (cuuint32_t)gm.TTh, 1};
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

template <int C>
int run_gt_fwd(const CUtensorMap& tmZ, const CUtensorMap& tmW, const float* bn, const float* pack, float* out, float* lin_out,
               const GtGeom& gm, int tiles, float drop_p, uint64_t seed, const uint64_t* seed_dev, uint64_t drop_stream,
               cudaStream_t s) {
    static bool cfg = false;
    if (!cfg) {
        int rc = opt_in_smem(bnglu_tc5_fwd_kernel<C>, GT_SMEM_FWD);
        if (rc) return rc;
        cfg = true;
    }
    const int grid = tiles < num_sms() ? tiles : num_sms();
    SEDK_CUDA(pdl_launch(bnglu_tc5_fwd_kernel<C>, dim3(grid), dim3(GT_THREADS), (size_t)(GT_SMEM_FWD), s, tmZ, tmW, bn, pack + GT_PACK_B, out, lin_out, gm, tiles,
                                                                  drop_threshold16(drop_p),
                                                                  drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f, seed,
                                                                  seed_dev, drop_stream));
    SEDK_LAUNCH_CHECK("bnglu_tc5_fwd_kernel");
    return SEDK_OK;
}

template <int C>
int run_gt_bwd(const CUtensorMap& tmZ, const CUtensorMap& tmW, const float* bn, const float* gout, float* lin_glin, float* gy,
               float* gglu_b, double* stats, const GtGeom& gm, int tiles, float drop_p, uint64_t seed,
               const uint64_t* seed_dev, uint64_t drop_stream, cudaStream_t s) {
    static bool cfg = false;
    if (!cfg) {
        int rc = opt_in_smem(bnglu_tc5_bwd_kernel<C>, GT_SMEM_BWD);
        if (rc) return rc;
        cfg = true;
    }
    const int grid = tiles < num_sms() ? tiles : num_sms();
    SEDK_CUDA(pdl_launch(bnglu_tc5_bwd_kernel<C>, dim3(grid), dim3(GT_THREADS), (size_t)(GT_SMEM_BWD), s, tmZ, tmW, bn, gout, lin_glin, gy, gglu_b, stats, gm, tiles,
                                                                  drop_threshold16(drop_p),
                                                                  drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f, seed,
                                                                  seed_dev, drop_stream));
    SEDK_LAUNCH_CHECK("bnglu_tc5_bwd_kernel");
    return SEDK_OK;
}

}  // namespace

bool bnglu_tc5_supports(int T, int F, int C, int pt, int pf, int precision) {
    GtGeom g;
    return precision == 0 && tc5_enabled() && get_option("bnglu_tc5", 1) != 0 && make_gtgeom(g, T, F, C, pt, pf);
}

int bnglu_tc5_pack_floats() { return GT_PACK_RAW + GT_C * GT_C; }

int launch_glu_prep(const double* stats, const float* gamma, const float* beta, float* running_mean, float* running_var,
                    int64_t* num_batches, float* bn, const float* glu_w, const float* glu_b, float* pack, double count,
                    float eps, float momentum, int training, int C, cudaStream_t s) {
    SEDK_PROF("glu_prep", s);
    SEDK_REQUIRE(C == 64 || C == 128, "glu_prep: C must be 64 or 128");
    if (C == 128)
        SEDK_CUDA(pdl_launch(glu_prep_kernel<128>, dim3(GT_C), dim3(GT_C), (size_t)(0), s, stats, gamma, beta, running_mean, running_var, num_batches, bn, glu_w, glu_b,
                                                   pack, count, eps, momentum, training));
    else
        SEDK_CUDA(pdl_launch(glu_prep_kernel<64>, dim3(GT_C), dim3(GT_C), (size_t)(0), s, stats, gamma, beta, running_mean, running_var, num_batches, bn, glu_w, glu_b,
                                                  pack, count, eps, momentum, training));
    SEDK_LAUNCH_CHECK("glu_prep_kernel");
    return SEDK_OK;
}

int launch_bnglu_tc5_fwd(const float* z, const float* bn, const float* pack, float* out, float* lin_out, int B, int T, int F,
                         int C, int pt, int pf, float drop_p, uint64_t seed, const uint64_t* seed_dev, uint64_t drop_stream,
                         cudaStream_t s) {
    char pname[64];
    snprintf(pname, sizeof(pname), "bnglu_tc5_fwd_c%d", C);
    SEDK_PROF(pname, s);
    GtGeom gm;
    SEDK_REQUIRE(make_gtgeom(gm, T, F, C, pt, pf), "bnglu_tc5: unsupported geometry");
    CUtensorMap tmZ, tmW;
    int rc = make_maps(z, pack, B, C, gm, &tmZ, &tmW);
    if (rc) return rc;
    const int tiles = B * gm.nTt;
    if (C == 128)
        return run_gt_fwd<128>(tmZ, tmW, bn, pack, out, lin_out, gm, tiles, drop_p, seed, seed_dev, drop_stream, s);
    return run_gt_fwd<64>(tmZ, tmW, bn, pack, out, lin_out, gm, tiles, drop_p, seed, seed_dev, drop_stream, s);
}

int launch_bnglu_tc5_bwd(const float* z, const float* bn, const float* pack, const float* gout, float* lin_glin, float* gy,
                         float* gglu_b, double* stats, int B, int T, int F, int C, int pt, int pf, float drop_p, uint64_t seed,
                         const uint64_t* seed_dev, uint64_t drop_stream, cudaStream_t s) {
    char pname[64];
    snprintf(pname, sizeof(pname), "bnglu_tc5_bwd_c%d", C);
    SEDK_PROF(pname, s);
    GtGeom gm;
    SEDK_REQUIRE(make_gtgeom(gm, T, F, C, pt, pf), "bnglu_tc5: unsupported geometry");
    CUtensorMap tmZ, tmW;
    int rc = make_maps(z, pack + GT_C * GT_C, B, C, gm, &tmZ, &tmW);          // A operand = WT
    if (rc) return rc;
    const int tiles = B * gm.nTt;
    if (C == 128)
        return run_gt_bwd<128>(tmZ, tmW, bn, gout, lin_glin, gy, gglu_b, stats, gm, tiles, drop_p, seed, seed_dev, drop_stream, s);
    return run_gt_bwd<64>(tmZ, tmW, bn, gout, lin_glin, gy, gglu_b, stats, gm, tiles, drop_p, seed, seed_dev, drop_stream, s);
}

// gglu_w <- BatchNorm-fold fix-up of g_lin^T z (tcgen05 TN GEMM over all pixels); needs gglu_b = sum g_lin complete.
// C = 128: the product accumulates straight into gglu_w (which must be zero on entry); C = 64: the 64-channel tensors are
// viewed as [.., F / 2, 128] pixel pairs, the 128 x 128 product goes to the scratch area of `pack` (zeroed here).
int launch_glu_wgrad_tc5(const float* z, const float* g_lin, const float* bn, float* pack, float* gglu_w, const float* gglu_b,
                         int B, int T, int F, int C, cudaStream_t s) {
    float* raw = gglu_w;
    int Fv = F;
    if (C == 64) {
        raw = pack + GT_PACK_RAW;
        Fv = F / 2;
        SEDK_CUDA(cudaMemsetAsync(raw, 0, (size_t)GT_C * GT_C * sizeof(float), s));
    }
    int rc = launch_tn_gemm_tc5_c128(z, g_lin, raw, B, T, Fv, s);
    if (rc) return rc;
    SEDK_PROF("glu_wgrad_fix", s);
    if (C == 128) glu_wgrad_fix_kernel<128><<<128, 128, 0, s>>>(gglu_w, raw, gglu_b, bn);
    else glu_wgrad_fix_kernel<64><<<64, 64, 0, s>>>(gglu_w, raw, gglu_b, bn);
    SEDK_LAUNCH_CHECK("glu_wgrad_fix_kernel");
    return SEDK_OK;
}

}  // namespace sedk
