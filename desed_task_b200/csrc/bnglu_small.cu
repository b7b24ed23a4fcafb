// BatchNorm -> GLU -> Dropout -> AvgPool for the narrow layers (C = 16, 32), forward and backward, "warp-autonomous":
// no shared-memory tile, no block barrier in the main loop.  Reference: desed_task/nnet/CNN.py:5-16, :73-98.
//
// ncu on the tiled kernels (bnglu.cu) at C = 16 / 32: 8 warps per SM, 4 barriers per 128-pixel tile, 110 instructions per
// element -> 277 / 239 us per launch against an HBM floor of ~45 / 22 us.  For C <= 32 the gate weight (C x C) fits in
// registers as ready-made mma.sync B fragments, so a warp can own 16 pixels end to end:
//   * lane (g = lane >> 2, t4 = lane & 3) loads channels 16 q + 4 t4 .. + 3 of pixels g and g + 8 as float4 (coalesced);
//   * the K index of the gate GEMM is permuted so that those float4s ARE the A fragments (k-slot t4 <-> channel
//     16 q + 4 t4 + 2 s, k-slot t4 + 4 <-> the next channel), and the N index is permuted so that the D fragment of a lane
//     lands on exactly the channels it loaded: BN affine, sigmoid, dropout, pooling and the stores all stay in registers;
//   * backward: lin = y Wg^T, g_lin, the elementwise term and g_y = g_lin Wg chain in registers; only the weight-gradient
//     GEMM (K = pixels) needs a transposed view, staged through a 2 x 16 x C per-warp shared-memory patch (__syncwarp
//     only); its accumulators and the per-channel BatchNorm sums persist in registers across the warp's group loop.
// Dropout uses 16-bit draws (8 elements per Philox4x32-7 call); the forward and backward kernels of this file share the
// (group, lane, chunk) -> counter mapping, which is all that is required (masks are regenerated, never stored).
#include "bnglu_small.cuh"

namespace sedk {
namespace {

// =====================================================================================================================
template <int C, bool X3>
__global__ void __launch_bounds__(256)
bnglu_small_fwd_kernel(const float* __restrict__ z, const float* __restrict__ bn, const float* __restrict__ glu_w,
                       const float* __restrict__ glu_b, float* __restrict__ out, SGeom gm, uint32_t thresh16,
                       float inv_keep, uint64_t seed, const uint64_t* __restrict__ seed_dev, uint64_t dstream) {
    pdl_enter();
    constexpr int Q = C / 16, NF = 2 * Q;
    __shared__ __align__(16) float vec[3 * C];      // scale, shift, gate bias
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    for (int i = tid; i < C; i += 256) {
        vec[i] = bn[i];
        vec[C + i] = bn[C + i];
        vec[2 * C + i] = glu_b[i];
    }
    GateB<C, X3> B1;
    B1.template load<1>(glu_w, g, t4);
    const Philox ph(seed + (seed_dev ? *seed_dev : 0ull));
    const float inv_pool = gm.pt == 2 ? 0.25f : 0.5f;
    __syncthreads();

    const int nwarps = gridDim.x * 8;
    int grp = blockIdx.x * 8 + warp;
    float4 zn[2][Q];
    if (grp < gm.total) {
        const GroupPos p = group_pos(gm, grp);
#pragma unroll
        for (int rr = 0; rr < 2; rr++)
#pragma unroll
            for (int q = 0; q < Q; q++)
                zn[rr][q] = __ldg(reinterpret_cast<const float4*>(z + pix_off<C>(gm, p, rr, g) + 16 * q + 4 * t4));
    }
    for (; grp < gm.total; grp += nwarps) {
        const GroupPos p = group_pos(gm, grp);
        float y[2][Q][4];
#pragma unroll
        for (int q = 0; q < Q; q++) {
            const float4 sc = *reinterpret_cast<const float4*>(vec + 16 * q + 4 * t4);
            const float4 sh = *reinterpret_cast<const float4*>(vec + C + 16 * q + 4 * t4);
#pragma unroll
            for (int rr = 0; rr < 2; rr++)
#pragma unroll
                for (int e = 0; e < 4; e++) y[rr][q][e] = fmaf(f4get(zn[rr][q], e), f4get(sc, e), f4get(sh, e));
        }
        // software pipeline: the next group's z is in flight while this one computes
        if (grp + nwarps < gm.total) {
            const GroupPos pn = group_pos(gm, grp + nwarps);
#pragma unroll
            for (int rr = 0; rr < 2; rr++)
#pragma unroll
                for (int q = 0; q < Q; q++)
                    zn[rr][q] = __ldg(reinterpret_cast<const float4*>(z + pix_off<C>(gm, pn, rr, g) + 16 * q + 4 * t4));
        }
        float acc[NF][4];
#pragma unroll
        for (int q = 0; q < Q; q++) {
            const float4 bg = *reinterpret_cast<const float4*>(vec + 2 * C + 16 * q + 4 * t4);
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
                acc[2 * q][2 * rr] = bg.x; acc[2 * q][2 * rr + 1] = bg.y;
                acc[2 * q + 1][2 * rr] = bg.z; acc[2 * q + 1][2 * rr + 1] = bg.w;
            }
        }
        gate_gemm<C, X3, 1>(acc, y, B1, glu_w, g, t4);
        // (lin * sigmoid(y)) -> dropout, in place in y
#pragma unroll
        for (int q = 0; q < Q; q++) {
            uint32_t kb = 0xffu;
            if (thresh16 != 0u) kb = keep_bits(ph, grp, lane, Q, q, dstream, thresh16);
#pragma unroll
            for (int rr = 0; rr < 2; rr++)
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    float v = acc[2 * q + (e >> 1)][2 * rr + (e & 1)] * lean_sigmoidf(y[rr][q][e]);
                    if (thresh16 != 0u) v = ((kb >> (4 * rr + e)) & 1u) ? v * inv_keep : 0.f;
                    y[rr][q][e] = v;
                }
        }
        // average pooling: partner along the mel axis is lane ^ 4 (g ^ 1); along time (pt == 2) it is the lane's other pixel
        if (gm.pt == 2) {
#pragma unroll
            for (int q = 0; q < Q; q++) {
                float s[4];
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    s[e] = y[0][q][e] + y[1][q][e];
                    s[e] = (s[e] + __shfl_xor_sync(0xffffffffu, s[e], 4)) * inv_pool;
                    // TF32 mode: the pooled activation is the next convolution's MMA operand (the tcgen05 unit truncates)
                    if (!X3) s[e] = __uint_as_float(to_tf32(s[e]));
                }
                if ((g & 1) == 0)
                    *reinterpret_cast<float4*>(out + pool_off<C>(gm, p, 0, g) + 16 * q + 4 * t4) =
                        make_float4(s[0], s[1], s[2], s[3]);
            }
        } else {
#pragma unroll
            for (int rr = 0; rr < 2; rr++)
#pragma unroll
                for (int q = 0; q < Q; q++) {
                    float s[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        s[e] = (y[rr][q][e] + __shfl_xor_sync(0xffffffffu, y[rr][q][e], 4)) * inv_pool;
                        if (!X3) s[e] = __uint_as_float(to_tf32(s[e]));
                    }
                    if ((g & 1) == 0)
                        *reinterpret_cast<float4*>(out + pool_off<C>(gm, p, rr, g) + 16 * q + 4 * t4) =
                            make_float4(s[0], s[1], s[2], s[3]);
                }
        }
    }
}

// =====================================================================================================================
template <int C, bool X3>
__global__ void __launch_bounds__(256)
bnglu_small_bwd_kernel(const float* __restrict__ z, const float* __restrict__ bn, const float* __restrict__ glu_w,
                       const float* __restrict__ glu_b, const float* __restrict__ gout, float* __restrict__ gy,
                       float* __restrict__ gglu_w, float* __restrict__ gglu_b, double* __restrict__ stats, SGeom gm,
                       uint32_t thresh16, float inv_keep, uint64_t seed, const uint64_t* __restrict__ seed_dev,
                       uint64_t dstream) {
    pdl_enter();
    constexpr int Q = C / 16, NF = 2 * Q;
    constexpr int S = C + 8;                                  // row stride of the staging patch (bank-conflict-free reads)
    constexpr int PATCH = 2 * 16 * S;                         // floats per warp: g_lin[16][S], y[16][S]
    constexpr int STAGE = 8 * PATCH > C * C ? 8 * PATCH : C * C;
    __shared__ __align__(16) float stage[STAGE];
    __shared__ __align__(16) float vec[5 * C];                // scale, shift, gate bias, -mean*invstd, invstd
    __shared__ float red[3 * C];                              // sum g_y, sum g_y*zhat, sum g_lin
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    for (int i = tid; i < C; i += 256) {
        vec[i] = bn[i];
        vec[C + i] = bn[C + i];
        vec[2 * C + i] = glu_b[i];
        vec[3 * C + i] = -bn[2 * C + i] * bn[3 * C + i];
        vec[4 * C + i] = bn[3 * C + i];
        red[i] = red[C + i] = red[2 * C + i] = 0.f;
    }
    GateB<C, X3> B1, B2;
    B1.template load<1>(glu_w, g, t4);
    B2.template load<2>(glu_w, g, t4);
    const Philox ph(seed + (seed_dev ? *seed_dev : 0ull));
    const float inv_pool = gm.pt == 2 ? 0.25f : 0.5f;
    float* Gs = stage + warp * PATCH;
    float* Ys = Gs + 16 * S;
    float dacc[Q][NF][4];
#pragma unroll
    for (int i = 0; i < Q; i++)
#pragma unroll
        for (int j = 0; j < NF; j++)
#pragma unroll
            for (int c = 0; c < 4; c++) dacc[i][j][c] = 0.f;
    float s1[Q][4], s2[Q][4], sg[Q][4];
#pragma unroll
    for (int q = 0; q < Q; q++)
#pragma unroll
        for (int e = 0; e < 4; e++) s1[q][e] = s2[q][e] = sg[q][e] = 0.f;
    __syncthreads();

    const int nwarps = gridDim.x * 8;
    for (int grp = blockIdx.x * 8 + warp; grp < gm.total; grp += nwarps) {
        const GroupPos p = group_pos(gm, grp);
        float4 zr[2][Q], go[2][Q];
#pragma unroll
        for (int rr = 0; rr < 2; rr++)
#pragma unroll
            for (int q = 0; q < Q; q++) {
                zr[rr][q] = __ldg(reinterpret_cast<const float4*>(z + pix_off<C>(gm, p, rr, g) + 16 * q + 4 * t4));
                if (rr == 0 || gm.pt == 1)
                    go[rr][q] = __ldg(reinterpret_cast<const float4*>(gout + pool_off<C>(gm, p, rr, g) + 16 * q + 4 * t4));
                else
                    go[rr][q] = go[0][q];
            }
        // L2 prefetch of the next group's lines (one 128-byte line per lane is enough for C <= 32)
        if (grp + nwarps < gm.total) {
            const GroupPos pn = group_pos(gm, grp + nwarps);
            asm volatile("prefetch.global.L2 [%0];\n" ::"l"(z + pix_off<C>(gm, pn, lane >> 4, (lane >> 1) & 7) + 16 * (lane & 1) * (Q - 1)));
        }
        float y[2][Q][4];
#pragma unroll
        for (int q = 0; q < Q; q++) {
            const float4 sc = *reinterpret_cast<const float4*>(vec + 16 * q + 4 * t4);
            const float4 sh = *reinterpret_cast<const float4*>(vec + C + 16 * q + 4 * t4);
#pragma unroll
            for (int rr = 0; rr < 2; rr++)
#pragma unroll
                for (int e = 0; e < 4; e++) y[rr][q][e] = fmaf(f4get(zr[rr][q], e), f4get(sc, e), f4get(sh, e));
        }
        float acc[NF][4];
#pragma unroll
        for (int q = 0; q < Q; q++) {
            const float4 bg = *reinterpret_cast<const float4*>(vec + 2 * C + 16 * q + 4 * t4);
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
                acc[2 * q][2 * rr] = bg.x; acc[2 * q][2 * rr + 1] = bg.y;
                acc[2 * q + 1][2 * rr] = bg.z; acc[2 * q + 1][2 * rr + 1] = bg.w;
            }
        }
        gate_gemm<C, X3, 1>(acc, y, B1, glu_w, g, t4);          // lin
        float gl[2][Q][4];
#pragma unroll
        for (int q = 0; q < Q; q++) {
            uint32_t kb = 0xffu;
            if (thresh16 != 0u) kb = keep_bits(ph, grp, lane, Q, q, dstream, thresh16);
#pragma unroll
            for (int rr = 0; rr < 2; rr++)
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    float ga = f4get(go[rr][q], e) * inv_pool;
                    if (thresh16 != 0u) ga = ((kb >> (4 * rr + e)) & 1u) ? ga * inv_keep : 0.f;
                    const float sgm = lean_sigmoidf(y[rr][q][e]);
                    float& a = acc[2 * q + (e >> 1)][2 * rr + (e & 1)];
                    const float g_lin = ga * sgm;
                    gl[rr][q][e] = g_lin;
                    a = ga * a * sgm * (1.0f - sgm);                 // elementwise part of g_y; GEMM 2 accumulates on top
                    sg[q][e] += g_lin;
                }
        }
        // stage g_lin and y (pixel-major) for the weight-gradient GEMM
#pragma unroll
        for (int rr = 0; rr < 2; rr++)
#pragma unroll
            for (int q = 0; q < Q; q++) {
                *reinterpret_cast<float4*>(Gs + (g + 8 * rr) * S + 16 * q + 4 * t4) =
                    make_float4(gl[rr][q][0], gl[rr][q][1], gl[rr][q][2], gl[rr][q][3]);
                *reinterpret_cast<float4*>(Ys + (g + 8 * rr) * S + 16 * q + 4 * t4) =
                    make_float4(y[rr][q][0], y[rr][q][1], y[rr][q][2], y[rr][q][3]);
            }
        gate_gemm<C, X3, 2>(acc, gl, B2, glu_w, g, t4);         // g_y
#pragma unroll
        for (int q = 0; q < Q; q++) {
            const float4 mi = *reinterpret_cast<const float4*>(vec + 3 * C + 16 * q + 4 * t4);
            const float4 is = *reinterpret_cast<const float4*>(vec + 4 * C + 16 * q + 4 * t4);
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
                const float4 o = make_float4(acc[2 * q][2 * rr], acc[2 * q][2 * rr + 1], acc[2 * q + 1][2 * rr],
                                             acc[2 * q + 1][2 * rr + 1]);
                *reinterpret_cast<float4*>(gy + pix_off<C>(gm, p, rr, g) + 16 * q + 4 * t4) = o;
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const float v = f4get(o, e);
                    const float zh = fmaf(f4get(zr[rr][q], e), f4get(is, e), f4get(mi, e));
                    s1[q][e] += v;
                    s2[q][e] = fmaf(v, zh, s2[q][e]);
                }
            }
        }
        __syncwarp();
        // dWg[n][k] += sum_pix g_lin[pix][n] * y[pix][k]   (A = g_lin^T, B = y; K = the 16 pixels of the group)
#pragma unroll
        for (int ks = 0; ks < 2; ks++) {
            uint32_t bh[NF][2], bl[NF][2];
#pragma unroll
            for (int j = 0; j < NF; j++) {
                const float b0 = Ys[(8 * ks + t4) * S + 8 * j + g], b1 = Ys[(8 * ks + t4 + 4) * S + 8 * j + g];
                bh[j][0] = to_tf32(b0); bh[j][1] = to_tf32(b1);
                if (X3) {
                    bl[j][0] = to_tf32(b0 - __uint_as_float(bh[j][0]));
                    bl[j][1] = to_tf32(b1 - __uint_as_float(bh[j][1]));
                }
            }
#pragma unroll
            for (int i = 0; i < Q; i++) {
                const float a[4] = {Gs[(8 * ks + t4) * S + 16 * i + g], Gs[(8 * ks + t4) * S + 16 * i + g + 8],
                                    Gs[(8 * ks + t4 + 4) * S + 16 * i + g], Gs[(8 * ks + t4 + 4) * S + 16 * i + g + 8]};
                uint32_t ah[4], al[4];
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    ah[c] = to_tf32(a[c]);
                    if (X3) al[c] = to_tf32(a[c] - __uint_as_float(ah[c]));
                }
#pragma unroll
                for (int j = 0; j < NF; j++) {
                    if (X3) {
                        mma_tf32(dacc[i][j], al, bh[j]);
                        mma_tf32(dacc[i][j], ah, bl[j]);
                    }
                    mma_tf32(dacc[i][j], ah, bh[j]);
                }
            }
        }
        __syncwarp();
    }

    // ---- flush: per-channel sums (reduce over the 8 pixel rows g of the warp, then CTA, then global)
#pragma unroll
    for (int q = 0; q < Q; q++)
#pragma unroll
        for (int e = 0; e < 4; e++) {
            float a = s1[q][e], b = s2[q][e], c = sg[q][e];
#pragma unroll
            for (int o = 4; o <= 16; o <<= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, o);
                b += __shfl_xor_sync(0xffffffffu, b, o);
                c += __shfl_xor_sync(0xffffffffu, c, o);
            }
            if (g == 0) {
                const int ch = 16 * q + 4 * t4 + e;
                atomicAdd(&red[ch], a);
                atomicAdd(&red[C + ch], b);
                atomicAdd(&red[2 * C + ch], c);
            }
        }
    __syncthreads();                                             // every warp is done with its staging patch
    for (int i = tid; i < C * C; i += 256) stage[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < Q; i++)
#pragma unroll
        for (int j = 0; j < NF; j++)
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int n = 16 * i + g + 8 * (c >> 1), k = 8 * j + 2 * t4 + (c & 1);
                atomicAdd(&stage[n * C + k], dacc[i][j][c]);
            }
    __syncthreads();
    for (int i = tid; i < C * C; i += 256) atomicAdd(&gglu_w[i], stage[i]);
    for (int i = tid; i < C; i += 256) {
        atomicAdd(&stats[2 * C + i], (double)red[i]);
        atomicAdd(&stats[3 * C + i], (double)red[C + i]);
        atomicAdd(&gglu_b[i], red[2 * C + i]);
    }
}

template <int C>
int run_small_fwd(const float* z, const float* bn, const float* glu_w, const float* glu_b, float* out, const SGeom& gm,
                  float p, uint64_t seed, const uint64_t* seed_dev, uint64_t dstream, int precision, cudaStream_t s) {
    const uint32_t th = drop_threshold16(p);
    const float inv_keep = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
    if (precision) {
        auto k = bnglu_small_fwd_kernel<C, true>;
        static int grid_cap = 0;
        if (grid_cap == 0) grid_cap = small_grid(k, 1 << 30);
        const int need = cdiv(gm.total, 8);
        SEDK_CUDA(pdl_launch(k, dim3(grid_cap < need ? grid_cap : need), dim3(256), (size_t)(0), s, z, bn, glu_w, glu_b, out, gm, th, inv_keep, seed, seed_dev, dstream));
    } else {
        auto k = bnglu_small_fwd_kernel<C, false>;
        static int grid_cap = 0;
        if (grid_cap == 0) grid_cap = small_grid(k, 1 << 30);
        const int need = cdiv(gm.total, 8);
        SEDK_CUDA(pdl_launch(k, dim3(grid_cap < need ? grid_cap : need), dim3(256), (size_t)(0), s, z, bn, glu_w, glu_b, out, gm, th, inv_keep, seed, seed_dev, dstream));
    }
    SEDK_LAUNCH_CHECK("bnglu_small_fwd_kernel");
    return SEDK_OK;
}

template <int C>
int run_small_bwd(const float* z, const float* bn, const float* glu_w, const float* glu_b, const float* gout, float* gy,
                  float* gglu_w, float* gglu_b, double* stats, const SGeom& gm, float p, uint64_t seed,
                  const uint64_t* seed_dev, uint64_t dstream, int precision, cudaStream_t s) {
    const uint32_t th = drop_threshold16(p);
    const float inv_keep = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
    if (precision) {
        auto k = bnglu_small_bwd_kernel<C, true>;
        static int grid_cap = 0;
        if (grid_cap == 0) grid_cap = small_grid(k, 1 << 30);
        const int need = cdiv(gm.total, 8);
        SEDK_CUDA(pdl_launch(k, dim3(grid_cap < need ? grid_cap : need), dim3(256), (size_t)(0), s, z, bn, glu_w, glu_b, gout, gy, gglu_w, gglu_b, stats, gm, th,
                                                          inv_keep, seed, seed_dev, dstream));
    } else {
        auto k = bnglu_small_bwd_kernel<C, false>;
        static int grid_cap = 0;
        if (grid_cap == 0) grid_cap = small_grid(k, 1 << 30);
        const int need = cdiv(gm.total, 8);
        SEDK_CUDA(pdl_launch(k, dim3(grid_cap < need ? grid_cap : need), dim3(256), (size_t)(0), s, z, bn, glu_w, glu_b, gout, gy, gglu_w, gglu_b, stats, gm, th,
                                                          inv_keep, seed, seed_dev, dstream));
    }
    SEDK_LAUNCH_CHECK("bnglu_small_bwd_kernel");
    return SEDK_OK;
}

}  // namespace

bool bnglu_small_supports(int B, int T, int F, int C, int pt, int pf) {
    SGeom g;
    return (C == 16 || C == 32) && get_option("bnglu_small", 1) != 0 && make_sgeom(g, B, T, F, pt, pf);
}

int launch_bnglu_small_fwd(const float* z, const float* bn, const float* glu_w, const float* glu_b, float* out, int B,
                           int T, int F, int C, int pt, int pf, float drop_p, uint64_t seed, const uint64_t* seed_dev,
                           uint64_t drop_stream, int precision, cudaStream_t s) {
    SGeom gm;
    SEDK_REQUIRE(make_sgeom(gm, B, T, F, pt, pf), "bnglu_small: unsupported geometry");
    if (C == 16) return run_small_fwd<16>(z, bn, glu_w, glu_b, out, gm, drop_p, seed, seed_dev, drop_stream, precision, s);
    if (C == 32) return run_small_fwd<32>(z, bn, glu_w, glu_b, out, gm, drop_p, seed, seed_dev, drop_stream, precision, s);
    SEDK_UNSUPPORTED("bnglu_small: channel width %d not in {16, 32}", C);
}

int launch_bnglu_small_bwd(const float* z, const float* bn, const float* glu_w, const float* glu_b, const float* gout,
                           float* gy, float* gglu_w, float* gglu_b, double* stats, int B, int T, int F, int C, int pt,
                           int pf, float drop_p, uint64_t seed, const uint64_t* seed_dev, uint64_t drop_stream,
                           int precision, cudaStream_t s) {
    SGeom gm;
    SEDK_REQUIRE(make_sgeom(gm, B, T, F, pt, pf), "bnglu_small: unsupported geometry");
    if (C == 16)
        return run_small_bwd<16>(z, bn, glu_w, glu_b, gout, gy, gglu_w, gglu_b, stats, gm, drop_p, seed, seed_dev,
                                 drop_stream, precision, s);
    if (C == 32)
        return run_small_bwd<32>(z, bn, glu_w, glu_b, gout, gy, gglu_w, gglu_b, stats, gm, drop_p, seed, seed_dev,
                                 drop_stream, precision, s);
    SEDK_UNSUPPORTED("bnglu_small: channel width %d not in {16, 32}", C);
}

}  // namespace sedk
