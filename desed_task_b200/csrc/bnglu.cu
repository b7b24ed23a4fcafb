// BatchNorm2d(eps 1e-3, momentum 0.99) -> GLU (learned 1x1 gate) -> Dropout -> AvgPool2d, fused, forward and backward.
// Reference: desed_task/nnet/CNN.py:5-16 (GLU = Linear_{C->C}(x) * sigmoid(x) over the channel axis), :73-98.
//
// forward  : one pass over the conv output z: y = scale*z + shift is formed while the 128-pixel tile is staged in shared
//            memory, the 1x1 gate is a tensor-core GEMM against the resident [C x C] weight, the epilogue applies
//            sigmoid*mul + Philox dropout in registers, and the pooled tile is written once (channels-last).
// backward : recomputes y / lin from z (nothing but z is saved), then in ONE kernel: pool-bwd + dropout-bwd + gate-bwd,
//            the data GEMM g_y = g_lin Wg (+ elementwise term), the weight GEMM dWg = g_lin^T y (TN, register
//            accumulators across a persistent tile loop) and the per-channel reductions BatchNorm's backward needs.
#include "kernels.h"

namespace sedk {
namespace {

__global__ void bn_finalize_kernel(const double* __restrict__ stats, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ rm, float* __restrict__ rv,
                                   int64_t* __restrict__ nb, float* __restrict__ bn, double count, float eps,
                                   float momentum, int training, int C) {
    pdl_enter();
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) {
        float mean, invstd;
        if (training) {
            double m = stats[c] / count;
            double var = stats[C + c] / count - m * m;
            if (var < 0.0) var = 0.0;
            mean = (float)m;
            invstd = (float)(1.0 / sqrt(var + (double)eps));
            double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
            rm[c] = (1.0f - momentum) * rm[c] + momentum * mean;
            rv[c] = (1.0f - momentum) * rv[c] + momentum * (float)unbiased;
        } else {
            mean = rm[c];
            invstd = 1.0f / sqrtf(rv[c] + eps);
        }
        float scale = gamma[c] * invstd;
        bn[c] = scale;
        bn[C + c] = beta[c] - mean * scale;
        bn[2 * C + c] = mean;
        bn[3 * C + c] = invstd;
    }
    if (c == 0 && training && nb != nullptr) *nb += 1;
}

template <int C>
struct GluCfg {
    static constexpr int YS = C + 4;
    static constexpr int WM = C >= 64 ? 4 : 8, WN = 8 / WM;
    static constexpr int MF = (128 / WM) / 16, NF = (C / WN) / 8;
    // weight-gradient (TN) warp layout: output split DWM x DWN, pixel (K) split DWK
    static constexpr int DWM = C >= 64 ? 4 : (C == 32 ? 2 : 1);
    static constexpr int DWN = C >= 32 ? 2 : 1;
    static constexpr int DWK = 8 / (DWM * DWN);
    static constexpr int DMF = C / DWM / 16, DNF = C / DWN / 8;
    static constexpr size_t SMEM_FWD = (size_t)(128 * YS + C * YS + 3 * C) * sizeof(float);
    static constexpr size_t SMEM_BWD = (size_t)(2 * 128 * YS + C * YS + 8 * C) * sizeof(float);
};

struct TileGeom {
    int T, F, Te, Fe, pt, pf, TT, TF, nTt, nTf, To, Fo;
    int tf_shift, otw_shift;      // log2(TF), log2(TF / pf): tile sides are powers of two
};

// Dropout keep-flags of the 4 elements a thread owns in one pair of adjacent n-fragments (columns n0+2t, n0+2t+1 of
// fragment 2q and of fragment 2q+1) of pixel `pix`: one Philox call serves all four.  The element -> random mapping only has
// to agree between the forward and the backward kernel of the same layer, which share this tiling.
__device__ __forceinline__ uint4 philox_frag_pair(const Philox& ph, uint64_t pix, int pairs_per_pixel, int pair, int t4,
                                                  uint64_t stream) {
    return ph((pix * (uint64_t)pairs_per_pixel + (uint64_t)pair) * 4ull + (uint64_t)t4, stream);
}

// L2 prefetch of the z rows (and, backward, the pooled gout rows) of the tile this CTA will process NEXT: the persistent
// loop is otherwise synchronous (load tile -> compute), which exposes HBM latency once per tile.
template <int C>
__device__ __forceinline__ void prefetch_tile(const float* __restrict__ z, const float* __restrict__ gout,
                                              const TileGeom& gm, int tile, int total_tiles, int tid) {
    if (tile >= total_tiles) return;
    int r = tile;
    const int b = r / (gm.nTt * gm.nTf);
    r -= b * gm.nTt * gm.nTf;
    const int t0 = (r / gm.nTf) * gm.TT, f0 = (r % gm.nTf) * gm.TF;
    constexpr int LINES = 128 * C / 32;            // 128-byte lines of a 128-pixel tile
    for (int i = tid; i < LINES; i += 256) {
        const int p = (i * 32) / C, ch = (i * 32) % C;
        const int ty = p >> gm.tf_shift, tx = p & (gm.TF - 1);
        const int t = t0 + ty, f = f0 + tx;
        if (t < gm.Te && f < gm.Fe) {
            const float* a = z + (((size_t)b * gm.T + t) * gm.F + f) * C + ch;
            asm volatile("prefetch.global.L2 [%0];\n" ::"l"(a));
            if (gout != nullptr && (t % gm.pt) == 0 && (f % gm.pf) == 0) {
                const float* g = gout + (((size_t)b * gm.To + t / gm.pt) * gm.Fo + f / gm.pf) * C + ch;
                asm volatile("prefetch.global.L2 [%0];\n" ::"l"(g));
            }
        }
    }
}

template <int C, bool X3>
__global__ void __launch_bounds__(256)
bnglu_fwd_kernel(const float* __restrict__ z, const float* __restrict__ bn, const float* __restrict__ glu_w,
                 const float* __restrict__ glu_b, float* __restrict__ out, TileGeom gm, uint32_t thresh, float inv_keep,
                 uint64_t seed, const uint64_t* __restrict__ seed_dev, uint64_t dstream, int total_tiles, int act) {
    // act (CNN.py:81-88): 0 GLU (Wy + b) * sigmoid(y); 1 ContextGating y * sigmoid(Wy + b); 2 ReLU; 3 LeakyReLU(0.2)
    using Cfg = GluCfg<C>;
    constexpr int YS = Cfg::YS, WN = Cfg::WN, MF = Cfg::MF, NF = Cfg::NF;
    extern __shared__ float smem[];
    float* Y = smem;                 // [128][YS]
    float* W = Y + 128 * YS;         // [C][YS]
    float* vec = W + C * YS;         // scale, shift, bg
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int wm0 = (warp / WN) * (MF * 16), wn0 = (warp % WN) * (NF * 8);
    for (int idx = tid; idx < C * (C / 4); idx += 256) {
        int n = idx / (C / 4), q = idx - n * (C / 4);
        *reinterpret_cast<float4*>(W + n * YS + q * 4) =
            glu_w != nullptr ? *reinterpret_cast<const float4*>(glu_w + n * C + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int i = tid; i < C; i += 256) {
        vec[i] = bn[i];
        vec[C + i] = bn[C + i];
        vec[2 * C + i] = glu_b != nullptr ? glu_b[i] : 0.f;
    }
    const Philox ph(seed + (seed_dev ? *seed_dev : 0ull));
    const float inv_pool = 1.0f / (float)(gm.pt * gm.pf);
    const int TT = gm.TT, TF = gm.TF;
    __syncthreads();

    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int r = tile;
        const int b = r / (gm.nTt * gm.nTf);
        r -= b * gm.nTt * gm.nTf;
        const int t0 = (r / gm.nTf) * TT, f0 = (r % gm.nTf) * TF;
        for (int idx = tid; idx < 128 * (C / 4); idx += 256) {
            int p = idx / (C / 4), q = idx - p * (C / 4);
            int ty = p >> gm.tf_shift, tx = p & (TF - 1);
            int t = t0 + ty, f = f0 + tx;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t < gm.Te && f < gm.Fe) {
                float4 zz = *reinterpret_cast<const float4*>(z + (((size_t)b * gm.T + t) * gm.F + f) * C + q * 4);
                const float* sc = vec + q * 4;
                const float* sh = vec + C + q * 4;
                v.x = fmaf(zz.x, sc[0], sh[0]); v.y = fmaf(zz.y, sc[1], sh[1]);
                v.z = fmaf(zz.z, sc[2], sh[2]); v.w = fmaf(zz.w, sc[3], sh[3]);
            }
            *reinterpret_cast<float4*>(Y + p * YS + q * 4) = v;
        }
        prefetch_tile<C>(z, nullptr, gm, tile + gridDim.x, total_tiles, tid);
        __syncthreads();
        float acc[MF][NF][4];
#pragma unroll
        for (int i = 0; i < MF; i++)
#pragma unroll
            for (int j = 0; j < NF; j++)
#pragma unroll
                for (int q = 0; q < 4; q++) acc[i][j][q] = 0.f;
        {
            const uint32_t ya = smem_u32(Y) + 4u * (uint32_t)((wm0 + (lane & 7) + 8 * ((lane >> 3) & 1)) * YS + 4 * (lane >> 4));
            const uint32_t wb = smem_u32(W) + 4u * (uint32_t)((wn0 + (lane >> 4) * 8 + (lane & 7)) * YS + 4 * ((lane >> 3) & 1));
#pragma unroll 4
            for (int k8 = 0; k8 < C / 8; k8++) {
                auto fa = [&](int i) { return ya + 4u * (uint32_t)(i * 16 * YS + k8 * 8); };
                auto fb = [&](int jp) { return wb + 4u * (uint32_t)(jp * 16 * YS + k8 * 8); };
                warp_mma_k8_ldsm<MF, NF, X3>(acc, fa, fb);
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < MF; i++)
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
                const int m = wm0 + i * 16 + g + 8 * rr;
                const int ty = m >> gm.tf_shift, tx = m & (TF - 1);
                const size_t pix = ((size_t)b * gm.T + (t0 + ty)) * gm.F + (f0 + tx);
                uint4 rnd = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                for (int j = 0; j < NF; j++) {
                    const int n = wn0 + j * 8 + 2 * t4;
                    float2 y = *reinterpret_cast<float2*>(Y + m * YS + n);
                    const float l0 = acc[i][j][2 * rr] + vec[2 * C + n], l1 = acc[i][j][2 * rr + 1] + vec[2 * C + n + 1];
                    float a0, a1;
                    if (act == 0) {
                        a0 = l0 * fast_sigmoidf_(y.x);
                        a1 = l1 * fast_sigmoidf_(y.y);
                    } else if (act == 1) {
                        a0 = y.x * fast_sigmoidf_(l0);
                        a1 = y.y * fast_sigmoidf_(l1);
                    } else {
                        const float slope = act == 2 ? 0.f : 0.2f;
                        a0 = y.x > 0.f ? y.x : slope * y.x;
                        a1 = y.y > 0.f ? y.y : slope * y.y;
                    }
                    if (thresh != 0u) {
                        if ((j & 1) == 0) rnd = philox_frag_pair(ph, pix, C / 16, (wn0 >> 4) + (j >> 1), t4, dstream);
                        const bool k0 = ((j & 1) ? rnd.z : rnd.x) >= thresh, k1 = ((j & 1) ? rnd.w : rnd.y) >= thresh;
                        a0 = k0 ? a0 * inv_keep : 0.f;
                        a1 = k1 ? a1 * inv_keep : 0.f;
                    }
                    *reinterpret_cast<float2*>(Y + m * YS + n) = make_float2(a0, a1);
                }
            }
        __syncthreads();
        const int otw = TF / gm.pf, oth = TT / gm.pt;
        for (int idx = tid; idx < oth * otw * (C / 4); idx += 256) {
            int op = idx / (C / 4), q = idx - op * (C / 4);
            int oy = op >> gm.otw_shift, ox = op & (otw - 1);
            int to = t0 / gm.pt + oy, fo = f0 / gm.pf + ox;
            if (to < gm.To && fo < gm.Fo) {
                float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int wy = 0; wy < gm.pt; wy++)
                    for (int wx = 0; wx < gm.pf; wx++) {
                        float4 v = *reinterpret_cast<float4*>(Y + ((oy * gm.pt + wy) * TF + ox * gm.pf + wx) * YS + q * 4);
                        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
                    }
                s.x *= inv_pool; s.y *= inv_pool; s.z *= inv_pool; s.w *= inv_pool;
                if (!X3) {      // TF32 mode: the pooled activation is the next convolution's MMA operand (tcgen05 truncates)
                    s.x = __uint_as_float(to_tf32(s.x)); s.y = __uint_as_float(to_tf32(s.y));
                    s.z = __uint_as_float(to_tf32(s.z)); s.w = __uint_as_float(to_tf32(s.w));
                }
                *reinterpret_cast<float4*>(out + (((size_t)b * gm.To + to) * gm.Fo + fo) * C + q * 4) = s;
            }
        }
        __syncthreads();
    }
}

template <int C, bool X3>
__global__ void __launch_bounds__(256, 1)
bnglu_bwd_kernel(const float* __restrict__ z, const float* __restrict__ bn, const float* __restrict__ glu_w,
                 const float* __restrict__ glu_b, const float* __restrict__ gout, float* __restrict__ gy,
                 float* __restrict__ gglu_w, float* __restrict__ gglu_b, double* __restrict__ stats, TileGeom gm,
                 uint32_t thresh, float inv_keep, uint64_t seed, const uint64_t* __restrict__ seed_dev, uint64_t dstream,
                 int total_tiles, int act) {
    using Cfg = GluCfg<C>;
    constexpr int YS = Cfg::YS, WN = Cfg::WN, MF = Cfg::MF, NF = Cfg::NF;
    constexpr int DWM = Cfg::DWM, DWN = Cfg::DWN, DWK = Cfg::DWK, DMF = Cfg::DMF, DNF = Cfg::DNF;
    constexpr int KWD = 128 / DWK;
    extern __shared__ float smem[];
    float* Y = smem;                 // [128][YS]   y = BN output
    float* G = Y + 128 * YS;         // [128][YS]   g_lin
    float* W = G + 128 * YS;         // [C][YS]
    float* vec = W + C * YS;         // scale, shift, bg, mean, invstd | red: sum gy, sum gy*zhat, sum g_lin
    float* red = vec + 5 * C;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int wm0 = (warp / WN) * (MF * 16), wn0 = (warp % WN) * (NF * 8);
    const int dk = warp / (DWM * DWN), dm0 = ((warp / DWN) % DWM) * (DMF * 16), dn0 = (warp % DWN) * (DNF * 8);
    for (int idx = tid; idx < C * (C / 4); idx += 256) {
        int n = idx / (C / 4), q = idx - n * (C / 4);
        *reinterpret_cast<float4*>(W + n * YS + q * 4) =
            glu_w != nullptr ? *reinterpret_cast<const float4*>(glu_w + n * C + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int i = tid; i < C; i += 256) {
        vec[i] = bn[i];
        vec[C + i] = bn[C + i];
        vec[2 * C + i] = glu_b != nullptr ? glu_b[i] : 0.f;
        vec[3 * C + i] = bn[2 * C + i];
        vec[4 * C + i] = bn[3 * C + i];
        red[i] = red[C + i] = red[2 * C + i] = 0.f;
    }
    float dacc[DMF][DNF][4];
#pragma unroll
    for (int i = 0; i < DMF; i++)
#pragma unroll
        for (int j = 0; j < DNF; j++)
#pragma unroll
            for (int q = 0; q < 4; q++) dacc[i][j][q] = 0.f;
    const Philox ph(seed + (seed_dev ? *seed_dev : 0ull));
    const float inv_pool = 1.0f / (float)(gm.pt * gm.pf);
    const int TT = gm.TT, TF = gm.TF;
    __syncthreads();

    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int r = tile;
        const int b = r / (gm.nTt * gm.nTf);
        r -= b * gm.nTt * gm.nTf;
        const int t0 = (r / gm.nTf) * TT, f0 = (r % gm.nTf) * TF;
        for (int idx = tid; idx < 128 * (C / 4); idx += 256) {
            int p = idx / (C / 4), q = idx - p * (C / 4);
            int ty = p >> gm.tf_shift, tx = p & (TF - 1);
            int t = t0 + ty, f = f0 + tx;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t < gm.Te && f < gm.Fe) {
                float4 zz = *reinterpret_cast<const float4*>(z + (((size_t)b * gm.T + t) * gm.F + f) * C + q * 4);
                const float* sc = vec + q * 4;
                const float* sh = vec + C + q * 4;
                v.x = fmaf(zz.x, sc[0], sh[0]); v.y = fmaf(zz.y, sc[1], sh[1]);
                v.z = fmaf(zz.z, sc[2], sh[2]); v.w = fmaf(zz.w, sc[3], sh[3]);
            }
            *reinterpret_cast<float4*>(Y + p * YS + q * 4) = v;
        }
        prefetch_tile<C>(z, gout, gm, tile + gridDim.x, total_tiles, tid);
        __syncthreads();
        // ---- GEMM 1: lin = y Wg^T
        float acc[MF][NF][4];
#pragma unroll
        for (int i = 0; i < MF; i++)
#pragma unroll
            for (int j = 0; j < NF; j++)
#pragma unroll
                for (int q = 0; q < 4; q++) acc[i][j][q] = 0.f;
        {
            const uint32_t ya = smem_u32(Y) + 4u * (uint32_t)((wm0 + (lane & 7) + 8 * ((lane >> 3) & 1)) * YS + 4 * (lane >> 4));
            const uint32_t wb = smem_u32(W) + 4u * (uint32_t)((wn0 + (lane >> 4) * 8 + (lane & 7)) * YS + 4 * ((lane >> 3) & 1));
#pragma unroll 4
            for (int k8 = 0; k8 < C / 8; k8++) {
                auto fa = [&](int i) { return ya + 4u * (uint32_t)(i * 16 * YS + k8 * 8); };
                auto fb = [&](int jp) { return wb + 4u * (uint32_t)(jp * 16 * YS + k8 * 8); };
                warp_mma_k8_ldsm<MF, NF, X3>(acc, fa, fb);
            }
        }
        // ---- epilogue 1: g_lin -> G, elementwise part of g_y -> acc
        float cs[NF][2];
#pragma unroll
        for (int j = 0; j < NF; j++) cs[j][0] = cs[j][1] = 0.f;
#pragma unroll
        for (int i = 0; i < MF; i++)
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
                const int m = wm0 + i * 16 + g + 8 * rr;
                const int ty = m >> gm.tf_shift, tx = m & (TF - 1);
                const int t = t0 + ty, f = f0 + tx;
                const bool valid = (t < gm.Te) && (f < gm.Fe);
                const size_t pix = ((size_t)b * gm.T + t) * gm.F + f;
                const float* gop = gout + (((size_t)b * gm.To + (t >> (gm.pt - 1))) * gm.Fo + (f >> (gm.pf - 1))) * C;
                uint4 rnd = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                for (int j = 0; j < NF; j++) {
                    const int n = wn0 + j * 8 + 2 * t4;
                    float ga0 = 0.f, ga1 = 0.f;
                    if (valid) {
                        float2 go = *reinterpret_cast<const float2*>(gop + n);
                        ga0 = go.x * inv_pool;
                        ga1 = go.y * inv_pool;
                        if (thresh != 0u) {
                            if ((j & 1) == 0) rnd = philox_frag_pair(ph, pix, C / 16, (wn0 >> 4) + (j >> 1), t4, dstream);
                            const bool k0 = ((j & 1) ? rnd.z : rnd.x) >= thresh, k1 = ((j & 1) ? rnd.w : rnd.y) >= thresh;
                            ga0 = k0 ? ga0 * inv_keep : 0.f;
                            ga1 = k1 ? ga1 * inv_keep : 0.f;
                        }
                    }
                    float2 y = *reinterpret_cast<float2*>(Y + m * YS + n);
                    float l0 = acc[i][j][2 * rr] + vec[2 * C + n], l1 = acc[i][j][2 * rr + 1] + vec[2 * C + n + 1];
                    float gl0, gl1, e0, e1;         // gradient wrt the gate pre-activation / element-wise part of g_y
                    if (act == 0) {                 // out = l * sigmoid(y)
                        const float s0 = fast_sigmoidf_(y.x), s1 = fast_sigmoidf_(y.y);
                        gl0 = ga0 * s0; gl1 = ga1 * s1;
                        e0 = ga0 * l0 * s0 * (1.0f - s0); e1 = ga1 * l1 * s1 * (1.0f - s1);
                    } else if (act == 1) {          // out = y * sigmoid(l)
                        const float s0 = fast_sigmoidf_(l0), s1 = fast_sigmoidf_(l1);
                        gl0 = ga0 * y.x * s0 * (1.0f - s0); gl1 = ga1 * y.y * s1 * (1.0f - s1);
                        e0 = ga0 * s0; e1 = ga1 * s1;
                    } else {                        // (leaky) ReLU: no gate
                        const float slope = act == 2 ? 0.f : 0.2f;
                        gl0 = gl1 = 0.f;
                        e0 = y.x > 0.f ? ga0 : slope * ga0; e1 = y.y > 0.f ? ga1 : slope * ga1;
                    }
                    *reinterpret_cast<float2*>(G + m * YS + n) = make_float2(gl0, gl1);
                    acc[i][j][2 * rr] = e0;
                    acc[i][j][2 * rr + 1] = e1;
                    cs[j][0] += gl0;
                    cs[j][1] += gl1;
                }
            }
#pragma unroll
        for (int j = 0; j < NF; j++)
#pragma unroll
            for (int q = 0; q < 2; q++) {
#pragma unroll
                for (int o = 4; o <= 16; o <<= 1) cs[j][q] += __shfl_xor_sync(0xffffffffu, cs[j][q], o);
                if (g == 0) atomicAdd(&red[2 * C + wn0 + j * 8 + 2 * t4 + q], cs[j][q]);
            }
        __syncthreads();
        // ---- GEMM 2: g_y += g_lin Wg   (B[k = gate-out n][col = gate-in k] = W[n][k])
#pragma unroll 4
        for (int k8 = 0; k8 < C / 8; k8++) {
            auto fa = [&](int i, int rr, int c) { return G[(wm0 + i * 16 + g + 8 * rr) * YS + k8 * 8 + t4 + 4 * c]; };
            auto fb = [&](int j, int c) { return W[(k8 * 8 + t4 + 4 * c) * YS + wn0 + j * 8 + g]; };
            warp_mma_k8<MF, NF, X3>(acc, fa, fb);
        }
        // ---- GEMM 3 (TN): dWg[n][k] += sum_pix g_lin[pix][n] * y[pix][k]
#pragma unroll 2
        for (int k8 = 0; k8 < KWD / 8; k8++) {
            const int kb = dk * KWD + k8 * 8;
            auto fa = [&](int i, int rr, int c) { return G[(kb + t4 + 4 * c) * YS + dm0 + i * 16 + g + 8 * rr]; };
            auto fb = [&](int j, int c) { return Y[(kb + t4 + 4 * c) * YS + dn0 + j * 8 + g]; };
            warp_mma_k8<DMF, DNF, X3>(dacc, fa, fb);
        }
        // ---- epilogue 2: store g_y, reduce sum g_y and sum g_y * zhat
        float s1v[NF][2], s2v[NF][2];
#pragma unroll
        for (int j = 0; j < NF; j++) s1v[j][0] = s1v[j][1] = s2v[j][0] = s2v[j][1] = 0.f;
#pragma unroll
        for (int i = 0; i < MF; i++)
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
                const int m = wm0 + i * 16 + g + 8 * rr;
                const int ty = m >> gm.tf_shift, tx = m & (TF - 1);
                const int t = t0 + ty, f = f0 + tx;
                if (t < gm.Te && f < gm.Fe) {
                    const size_t pix = ((size_t)b * gm.T + t) * gm.F + f;
#pragma unroll
                    for (int j = 0; j < NF; j++) {
                        const int k = wn0 + j * 8 + 2 * t4;
                        float v0 = acc[i][j][2 * rr], v1 = acc[i][j][2 * rr + 1];
                        *reinterpret_cast<float2*>(gy + pix * C + k) = make_float2(v0, v1);
                        float2 zz = *reinterpret_cast<const float2*>(z + pix * C + k);
                        float zh0 = (zz.x - vec[3 * C + k]) * vec[4 * C + k];
                        float zh1 = (zz.y - vec[3 * C + k + 1]) * vec[4 * C + k + 1];
                        s1v[j][0] += v0; s1v[j][1] += v1;
                        s2v[j][0] = fmaf(v0, zh0, s2v[j][0]); s2v[j][1] = fmaf(v1, zh1, s2v[j][1]);
                    }
                }
            }
#pragma unroll
        for (int j = 0; j < NF; j++)
#pragma unroll
            for (int q = 0; q < 2; q++) {
#pragma unroll
                for (int o = 4; o <= 16; o <<= 1) {
                    s1v[j][q] += __shfl_xor_sync(0xffffffffu, s1v[j][q], o);
                    s2v[j][q] += __shfl_xor_sync(0xffffffffu, s2v[j][q], o);
                }
                if (g == 0) {
                    atomicAdd(&red[wn0 + j * 8 + 2 * t4 + q], s1v[j][q]);
                    atomicAdd(&red[C + wn0 + j * 8 + 2 * t4 + q], s2v[j][q]);
                }
            }
        __syncthreads();
    }
    // ---- flush the persistent accumulators
    if (gglu_w != nullptr)
#pragma unroll
    for (int i = 0; i < DMF; i++)
#pragma unroll
        for (int j = 0; j < DNF; j++)
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
                const int n = dm0 + i * 16 + g + 8 * rr, k = dn0 + j * 8 + 2 * t4;
                atomicAdd(reinterpret_cast<float2*>(gglu_w + (size_t)n * C + k),
                          make_float2(dacc[i][j][2 * rr], dacc[i][j][2 * rr + 1]));
            }
    __syncthreads();
    for (int i = tid; i < C; i += 256) {
        atomicAdd(&stats[2 * C + i], (double)red[i]);
        atomicAdd(&stats[3 * C + i], (double)red[C + i]);
        if (gglu_b != nullptr) atomicAdd(&gglu_b[i], red[2 * C + i]);
    }
}

// gz = scale * (gy - mean(gy) - zhat * mean(gy * zhat)), in place; block 0 also writes the BN / bias gradients
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(float* __restrict__ gy, const float* __restrict__ z, const float* __restrict__ bn,
                    const double* __restrict__ stats, float* __restrict__ ggamma, float* __restrict__ gbeta,
                    float* __restrict__ gb, float inv_count, int64_t n4, int C, int frozen) {
    pdl_enter();
    extern __shared__ float sv[];    // scale, mean, invstd, m1, m2
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        sv[c] = bn[c];
        sv[C + c] = bn[2 * C + c];
        sv[2 * C + c] = bn[3 * C + c];
        double s1 = stats[2 * C + c], s2 = stats[3 * C + c];
        // running-statistics (frozen / eval) BatchNorm is a fixed affine map: no mean terms in its backward
        sv[3 * C + c] = frozen ? 0.f : (float)(s1 * (double)inv_count);
        sv[4 * C + c] = frozen ? 0.f : (float)(s2 * (double)inv_count);
        if (blockIdx.x == 0) {
            gbeta[c] = (float)s1;
            ggamma[c] = (float)s2;
            // the conv bias cancels inside a batch-statistics BatchNorm; through a frozen one it sees scale * sum gy
            if (gb != nullptr) gb[c] = frozen ? bn[c] * (float)s1 : 0.f;
        }
    }
    __syncthreads();
    const int c4n = C / 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4n) * 4;
        float4 g4 = reinterpret_cast<float4*>(gy)[i];
        float4 z4 = reinterpret_cast<const float4*>(z)[i];
        float* gv = reinterpret_cast<float*>(&g4);
        const float* zv = reinterpret_cast<const float*>(&z4);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float zh = (zv[k] - sv[C + c + k]) * sv[2 * C + c + k];
            gv[k] = sv[c + k] * (gv[k] - sv[3 * C + c + k] - zh * sv[4 * C + c + k]);
        }
        reinterpret_cast<float4*>(gy)[i] = g4;
    }
}

inline TileGeom make_geom(int T, int F, int pt, int pf) {
    TileGeom g;
    g.T = T; g.F = F; g.pt = pt; g.pf = pf;
    g.To = T / pt; g.Fo = F / pf;
    g.Te = g.To * pt; g.Fe = g.Fo * pf;
    if (g.Fe > 8) { g.TT = 8; g.TF = 16; }
    else if (g.Fe > 4) { g.TT = 16; g.TF = 8; }
    else if (g.Fe > 2) { g.TT = 32; g.TF = 4; }
    else { g.TT = 64; g.TF = 2; }
    g.nTt = cdiv(g.Te, g.TT);
    g.nTf = cdiv(g.Fe, g.TF);
    g.tf_shift = 0;
    while ((1 << g.tf_shift) < g.TF) g.tf_shift++;
    g.otw_shift = 0;
    while ((1 << g.otw_shift) < g.TF / g.pf) g.otw_shift++;
    return g;
}

inline int geom_ok(const TileGeom& g) {
    if (g.To < 1 || g.Fo < 1) return 0;
    if (g.TT % g.pt != 0 || g.TF % g.pf != 0) return 0;
    return 1;
}

template <int C>
int run_fwd(const float* z, const float* bn, const float* glu_w, const float* glu_b, float* out, int B, TileGeom gm,
            float p, uint64_t seed, const uint64_t* seed_dev, uint64_t dstream, int precision, int act, cudaStream_t s) {
    using Cfg = GluCfg<C>;
    const int tiles = B * gm.nTt * gm.nTf;
    const uint32_t thresh = p > 0.f ? drop_threshold(p) : 0u;
    const float inv_keep = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
    static int occ[2] = {0, 0};
    const int pi = precision ? 1 : 0;
    auto k0 = bnglu_fwd_kernel<C, false>;
    auto k1 = bnglu_fwd_kernel<C, true>;
    if (occ[pi] == 0) {
        int rc = precision ? opt_in_smem(k1, Cfg::SMEM_FWD) : opt_in_smem(k0, Cfg::SMEM_FWD);
        if (rc) return rc;
        int o = 1;
        if (precision) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k1, 256, Cfg::SMEM_FWD);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k0, 256, Cfg::SMEM_FWD);
        occ[pi] = o < 1 ? 1 : o;
    }
    int grid = num_sms() * occ[pi];
    if (grid > tiles) grid = tiles;
    if (precision) k1<<<grid, 256, Cfg::SMEM_FWD, s>>>(z, bn, glu_w, glu_b, out, gm, thresh, inv_keep, seed, seed_dev, dstream, tiles, act);
    else k0<<<grid, 256, Cfg::SMEM_FWD, s>>>(z, bn, glu_w, glu_b, out, gm, thresh, inv_keep, seed, seed_dev, dstream, tiles, act);
    SEDK_LAUNCH_CHECK("bnglu_fwd_kernel");
    return SEDK_OK;
}

template <int C>
int run_bwd(const float* z, const float* bn, const float* glu_w, const float* glu_b, const float* gout, float* gy,
            float* gglu_w, float* gglu_b, double* stats, int B, TileGeom gm, float p, uint64_t seed, const uint64_t* seed_dev,
            uint64_t dstream, int precision, int act, cudaStream_t s) {
    using Cfg = GluCfg<C>;
    const int tiles = B * gm.nTt * gm.nTf;
    const uint32_t thresh = p > 0.f ? drop_threshold(p) : 0u;
    const float inv_keep = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
    static int occ[2] = {0, 0};
    const int pi = precision ? 1 : 0;
    auto k0 = bnglu_bwd_kernel<C, false>;
    auto k1 = bnglu_bwd_kernel<C, true>;
    if (occ[pi] == 0) {
        int rc = precision ? opt_in_smem(k1, Cfg::SMEM_BWD) : opt_in_smem(k0, Cfg::SMEM_BWD);
        if (rc) return rc;
        int o = 1;
        if (precision) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k1, 256, Cfg::SMEM_BWD);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k0, 256, Cfg::SMEM_BWD);
        occ[pi] = o < 1 ? 1 : o;
    }
    int grid = num_sms() * occ[pi];
    if (grid > tiles) grid = tiles;
    if (precision)
        k1<<<grid, 256, Cfg::SMEM_BWD, s>>>(z, bn, glu_w, glu_b, gout, gy, gglu_w, gglu_b, stats, gm, thresh, inv_keep, seed,
                                            seed_dev, dstream, tiles, act);
    else
        k0<<<grid, 256, Cfg::SMEM_BWD, s>>>(z, bn, glu_w, glu_b, gout, gy, gglu_w, gglu_b, stats, gm, thresh, inv_keep, seed,
                                            seed_dev, dstream, tiles, act);
    SEDK_LAUNCH_CHECK("bnglu_bwd_kernel");
    return SEDK_OK;
}

}  // namespace

int launch_bn_finalize(const double* stats, const float* gamma, const float* beta, float* running_mean,
                       float* running_var, int64_t* num_batches, float* bn, double count, float eps, float momentum,
                       int training, int C, cudaStream_t s) {
    SEDK_PROF("bn_finalize", s);
    SEDK_CUDA(pdl_launch(bn_finalize_kernel, dim3(cdiv(C, 128)), dim3(128), (size_t)(0), s, stats, gamma, beta, running_mean, running_var, num_batches, bn, count,
                                                    eps, momentum, training, C));
    SEDK_LAUNCH_CHECK("bn_finalize_kernel");
    return SEDK_OK;
}

int launch_bnglu_pool_fwd(const float* z, const float* bn, const float* glu_w, const float* glu_b, float* out, int B,
                          int T, int F, int C, int pt, int pf, float drop_p, uint64_t seed, const uint64_t* seed_dev,
                          uint64_t drop_stream, int precision, int act, cudaStream_t s) {
    char pname[64];
    snprintf(pname, sizeof(pname), "bnglu_pool_fwd_c%d", C);
    SEDK_PROF(pname, s);
    SEDK_REQUIRE(act >= 0 && act <= 3, "bnglu_pool: unknown activation %d", act);
    SEDK_REQUIRE(act >= 2 || (glu_w != nullptr && glu_b != nullptr), "bnglu_pool: gate parameters missing");
    if (act == 0 && bnglu_small_supports(B, T, F, C, pt, pf))
        return launch_bnglu_small_fwd(z, bn, glu_w, glu_b, out, B, T, F, C, pt, pf, drop_p, seed, seed_dev, drop_stream,
                                      precision, s);
    TileGeom gm = make_geom(T, F, pt, pf);
    SEDK_REQUIRE(geom_ok(gm), "bnglu_pool: pooling (%d,%d) on a %dx%d map is not supported", pt, pf, T, F);
    switch (C) {
        case 16: return run_fwd<16>(z, bn, glu_w, glu_b, out, B, gm, drop_p, seed, seed_dev, drop_stream, precision, act, s);
        case 32: return run_fwd<32>(z, bn, glu_w, glu_b, out, B, gm, drop_p, seed, seed_dev, drop_stream, precision, act, s);
        case 64: return run_fwd<64>(z, bn, glu_w, glu_b, out, B, gm, drop_p, seed, seed_dev, drop_stream, precision, act, s);
        case 128: return run_fwd<128>(z, bn, glu_w, glu_b, out, B, gm, drop_p, seed, seed_dev, drop_stream, precision, act, s);
    }
    SEDK_UNSUPPORTED("bnglu_pool: channel width %d not in {16,32,64,128}", C);
}

int launch_bnglu_pool_bwd(const float* z, const float* bn, const float* glu_w, const float* glu_b, const float* gout,
                          float* gy, float* gglu_w, float* gglu_b, double* stats, int B, int T, int F, int C, int pt,
                          int pf, float drop_p, uint64_t seed, const uint64_t* seed_dev, uint64_t drop_stream, int precision,
                          int act, cudaStream_t s) {
    char pname[64];
    snprintf(pname, sizeof(pname), "bnglu_pool_bwd_c%d", C);
    SEDK_PROF(pname, s);
    TileGeom gm = make_geom(T, F, pt, pf);
    SEDK_REQUIRE(geom_ok(gm), "bnglu_pool: pooling (%d,%d) on a %dx%d map is not supported", pt, pf, T, F);
    if (gm.Te != T || gm.Fe != F) SEDK_CUDA(cudaMemsetAsync(gy, 0, (size_t)B * T * F * C * sizeof(float), s));
    if (act == 0 && bnglu_small_supports(B, T, F, C, pt, pf))
        return launch_bnglu_small_bwd(z, bn, glu_w, glu_b, gout, gy, gglu_w, gglu_b, stats, B, T, F, C, pt, pf, drop_p, seed,
                                      seed_dev, drop_stream, precision, s);
    switch (C) {
        case 16: return run_bwd<16>(z, bn, glu_w, glu_b, gout, gy, gglu_w, gglu_b, stats, B, gm, drop_p, seed, seed_dev, drop_stream, precision, act, s);
        case 32: return run_bwd<32>(z, bn, glu_w, glu_b, gout, gy, gglu_w, gglu_b, stats, B, gm, drop_p, seed, seed_dev, drop_stream, precision, act, s);
        case 64: return run_bwd<64>(z, bn, glu_w, glu_b, gout, gy, gglu_w, gglu_b, stats, B, gm, drop_p, seed, seed_dev, drop_stream, precision, act, s);
        case 128: return run_bwd<128>(z, bn, glu_w, glu_b, gout, gy, gglu_w, gglu_b, stats, B, gm, drop_p, seed, seed_dev, drop_stream, precision, act, s);
    }
    SEDK_UNSUPPORTED("bnglu_pool: channel width %d not in {16,32,64,128}", C);
}

int launch_bn_bwd_apply(float* gy, const float* z, const float* bn, const double* stats, float* ggamma, float* gbeta,
                        float* gb, double count, int64_t n_pix, int C, int frozen, cudaStream_t s) {
    char pname[64];
    snprintf(pname, sizeof(pname), "bn_bwd_apply_c%d", C);
    SEDK_PROF(pname, s);
    SEDK_REQUIRE(C % 4 == 0, "bn_bwd_apply: C %% 4 != 0");
    const int64_t n4 = n_pix * C / 4;
    int64_t blocks = (n4 + 256 * 4 - 1) / (256 * 4);
    int64_t cap = (int64_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    SEDK_CUDA(pdl_launch(bn_bwd_apply_kernel, dim3((int)blocks), dim3(256), (size_t)(5 * C * sizeof(float)), s, gy, z, bn, stats, ggamma, gbeta, gb,
                                                                       (float)(1.0 / count), n4, C, frozen));
    SEDK_LAUNCH_CHECK("bn_bwd_apply_kernel");
    return SEDK_OK;
}

}  // namespace sedk
