// Shared pieces of the warp-autonomous narrow-layer kernels (bnglu_small.cu, layer0.cu): the 16-pixel group geometry, the
// permuted gate-GEMM fragments (see the layout note at the top of bnglu_small.cu) and the 16-bit dropout draws.
#pragma once
#include "kernels.h"

namespace sedk {
namespace {

struct SGeom {
    int T, F, To, Fo;
    int pt;        // 1 or 2 (pf is always 2 here)
    int gpr;       // 16-pixel groups per row: pt == 2 -> 2 rows x 8 bins (F / 8), pt == 1 -> 1 row x 16 bins (F / 16)
    int gpr_shift; // log2(gpr) when gpr is a power of two, else -1
    int rows;      // pt == 2 ? To : T
    float inv_rows;
    int total;     // B * rows * gpr
};

// MUFU-only sigmoid without the range-check code of __expf / __fdividef (ex2 -> inf gives rcp -> 0, the right limit)
__device__ __forceinline__ float lean_sigmoidf(float x) {
    float e, y;
    asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(e) : "f"(-1.4426950408889634f * x));
    asm("rcp.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(1.0f + e));
    return y;
}

// physical channel of n-slot `sigma` (0..7) of n-fragment j: chosen so that D fragment (cols 2 t4, 2 t4 + 1 of fragment
// j = 2 q + s) holds channels 16 q + 4 t4 + 2 s + {0, 1} - the same channels the lane loaded as float4 chunk q.
__device__ __forceinline__ int chan_slot(int j, int sigma) {
    return 16 * (j >> 1) + 4 * (sigma >> 1) + 2 * (j & 1) + (sigma & 1);
}

// B fragment of k-step (q, s) and n-fragment j.  WHICH 1: lin = y Wg^T (B[k][n] = W[n][k]);  2: g_y = g_lin Wg (B[k][n] = W[k][n])
template <int C, int WHICH>
__device__ __forceinline__ float2 gate_bfrag(const float* __restrict__ W, int q, int s, int j, int g, int t4) {
    const int k0 = 16 * q + 4 * t4 + 2 * s;
    const int n = chan_slot(j, g);
    if (WHICH == 1) return __ldg(reinterpret_cast<const float2*>(W + n * C + k0));
    return make_float2(__ldg(W + k0 * C + n), __ldg(W + (k0 + 1) * C + n));
}

template <int C, bool X3>
struct GateB {
    static constexpr int Q = C / 16, KS = 2 * Q, NF = 2 * Q;
    uint32_t h[X3 ? 1 : KS * NF][2];
    template <int WHICH>
    __device__ __forceinline__ void load(const float* __restrict__ W, int g, int t4) {
        if (!X3) {
#pragma unroll
            for (int ks = 0; ks < KS; ks++)
#pragma unroll
                for (int j = 0; j < NF; j++) {
                    const float2 w = gate_bfrag<C, WHICH>(W, ks >> 1, ks & 1, j, g, t4);
                    h[ks * NF + j][0] = to_tf32(w.x);
                    h[ks * NF + j][1] = to_tf32(w.y);
                }
        }
    }
};

// acc[j][2 rr + e] += sum_k v[rr][.][k] * B[k][chan(j, 2 t4 + e)]   (v in the float4-chunk register layout)
template <int C, bool X3, int WHICH>
__device__ __forceinline__ void gate_gemm(float (&acc)[C / 8][4], const float (&v)[2][C / 16][4], const GateB<C, X3>& B,
                                          const float* __restrict__ W, int g, int t4) {
    constexpr int Q = C / 16, NF = 2 * Q;
#pragma unroll
    for (int q = 0; q < Q; q++)
#pragma unroll
        for (int s = 0; s < 2; s++) {
            const float a[4] = {v[0][q][2 * s], v[1][q][2 * s], v[0][q][2 * s + 1], v[1][q][2 * s + 1]};
            uint32_t ah[4], al[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                // !X3: half-ulp add, the MMA unit drops the low bits; X3 needs the rounded value itself for the residual
                ah[i] = X3 ? to_tf32(a[i]) : to_tf32_mma(a[i]);
                if (X3) al[i] = to_tf32(a[i] - __uint_as_float(ah[i]));
            }
#pragma unroll
            for (int j = 0; j < NF; j++) {
                if (X3) {
                    const float2 w = gate_bfrag<C, WHICH>(W, q, s, j, g, t4);
                    uint32_t bh[2] = {to_tf32(w.x), to_tf32(w.y)};
                    uint32_t bl[2] = {to_tf32(w.x - __uint_as_float(bh[0])), to_tf32(w.y - __uint_as_float(bh[1]))};
                    mma_tf32(acc[j], al, bh);
                    mma_tf32(acc[j], ah, bl);
                    mma_tf32(acc[j], ah, bh);
                } else {
                    mma_tf32(acc[j], ah, B.h[(2 * q + s) * NF + j]);
                }
            }
        }
}

struct GroupPos {
    int b, trow, fg;
};
__device__ __forceinline__ GroupPos group_pos(const SGeom& gm, int grp) {
    GroupPos p;
    int r;
    if (gm.gpr_shift >= 0) {
        p.fg = grp & (gm.gpr - 1);
        r = grp >> gm.gpr_shift;
    } else {
        p.fg = grp % gm.gpr;
        r = grp / gm.gpr;
    }
    // r / rows through a float reciprocal: exact while r < 2^22 and rows < 2^11 (checked by make_sgeom)
    p.b = __float2int_rd(((float)r + 0.5f) * gm.inv_rows);
    p.trow = r - p.b * gm.rows;
    return p;
}
// element offset (in floats, without the channel term) of pixel (rr, g) of a group in a [B, T, F, C] tensor
template <int C>
__device__ __forceinline__ size_t pix_off(const SGeom& gm, const GroupPos& p, int rr, int g) {
    const int t = gm.pt == 2 ? 2 * p.trow + rr : p.trow;
    const int f = gm.pt == 2 ? 8 * p.fg + g : 16 * p.fg + 8 * rr + g;
    return (((size_t)p.b * gm.T + t) * gm.F + f) * C;
}
// pooled pixel that pixel (rr, g) contributes to, in a [B, To, Fo, C] tensor
template <int C>
__device__ __forceinline__ size_t pool_off(const SGeom& gm, const GroupPos& p, int rr, int g) {
    const int fo = gm.pt == 2 ? 4 * p.fg + (g >> 1) : 8 * p.fg + 4 * rr + (g >> 1);
    return (((size_t)p.b * gm.To + p.trow) * gm.Fo + fo) * C;
}

__device__ __forceinline__ float f4get(const float4& v, int e) { return e == 0 ? v.x : e == 1 ? v.y : e == 2 ? v.z : v.w; }

// keep flags of the 8 elements (2 pixels x 4 channels) of chunk q: bit (4 rr + e)
__device__ __forceinline__ uint32_t keep_bits(const Philox& ph, int grp, int lane, int Q, int q, uint64_t dstream,
                                              uint32_t thresh16) {
    const uint4 r = ph(((uint64_t)grp * 32ull + (uint64_t)lane) * (uint64_t)Q + (uint64_t)q, dstream);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    uint32_t bits = 0;
#pragma unroll
    for (int e = 0; e < 4; e++) {
        bits |= ((w[e] & 0xffffu) >= thresh16 ? 1u : 0u) << e;
        bits |= ((w[e] >> 16) >= thresh16 ? 1u : 0u) << (4 + e);
    }
    return bits;
}

inline bool make_sgeom(SGeom& g, int B, int T, int F, int pt, int pf) {
    if (pf != 2 || (pt != 1 && pt != 2)) return false;
    if (F % (pt == 2 ? 8 : 16) != 0 || T / pt < 1) return false;
    g.T = T; g.F = F; g.To = T / pt; g.Fo = F / 2; g.pt = pt;
    g.gpr = pt == 2 ? F / 8 : F / 16;
    g.rows = pt == 2 ? g.To : T;
    g.inv_rows = 1.0f / (float)g.rows;
    g.gpr_shift = -1;
    for (int sft = 0; sft < 16; sft++)
        if ((1 << sft) == g.gpr) g.gpr_shift = sft;
    const long long total = (long long)B * g.rows * g.gpr;
    if (total <= 0 || total > 0x7fffffffLL) return false;
    if ((long long)B * g.rows >= (1 << 22) || g.rows >= (1 << 11)) return false;    // float-reciprocal division range
    g.total = (int)total;
    return true;
}

inline uint32_t drop_threshold16(float p) {
    if (p <= 0.f) return 0u;
    double t = (double)p * 65536.0 + 0.5;
    if (t < 1.0) t = 1.0;
    if (t > 65535.0) t = 65535.0;
    return (uint32_t)t;
}

template <class K>
int small_grid(K kern, int groups) {
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 0);
    if (occ < 1) occ = 1;
    int grid = num_sms() * occ;
    const int need = cdiv(groups, 8);
    return grid < need ? grid : need;
}

}  // namespace
}  // namespace sedk
