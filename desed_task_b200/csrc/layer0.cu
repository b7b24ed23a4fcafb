// First CNN block with the convolution output never materialised (desed_task/nnet/CNN.py:66-98 for i = 0, n_in_channel = 1,
// 16 filters, GLU, pooling (2, 2); scaler utils/scaler.py:114-120 and SpecAugment CRNN.py:207-219 in front of it).
//
// The unfused path moves z0 = conv0(x) ([B, T, F, 16] fp32 = 123 MB at 24 clips) through HBM eight times per training step
// (conv0 write; BN+GLU forward read; backward: read z0, write g_y, BN-apply read g_y + z0 / write g_z, weight-gradient read
// g_z): 0.34 ms of the 2.19 ms step, all of it HBM time.  With ONE input channel z0 is a 9-tap stencil of a 7.7 MB image,
// cheaper to recompute (9 FMA per element) than to load, and the tail of the backward collapses into reductions because
// nothing upstream needs a data gradient:
//
//   l0_x0     x (strided log-mel) -> x0 [B, T, F] (scaled, masked) (+ border sums that give SX[tap] = sum_pix x0(pix + tap))
//   l0_stats  (batch-statistics forward only) the sums that depend on x alone:  sum z, sum z^2 (BatchNorm),
//             ZX[c][tap] = sum_pix z_c x0(pix + tap)
//   l0_fwd    x0 -> stencil -> BN -> gate GEMM -> sigmoid -> dropout -> 2x2 average pool -> out          (no z0 in HBM)
//   l0_bwd    g_out, x0 -> recompute up to the gate, g_y in registers -> S1 = sum g_y, S2 = sum g_y zhat,
//             GX[c][tap] = sum_pix g_y,c x0(pix + tap), gate weight / bias gradients                       (no g_y / g_z in HBM)
//   l0_finish BatchNorm backward folded into the weight gradient in closed form (fp64):
//               g_z = scale (g_y - S1/N - zhat S2/N),  zhat = (z - mean) invstd
//               dW[c][tap] = sum g_z x0(tap) = scale ( GX - (S1/N) SX - (S2/N) invstd (ZX - mean SX) )
//             (frozen BatchNorm: dW = scale GX, db = scale S1), plus ggamma = S2, gbeta = S1.
//
// The three big kernels are "strip walkers".  A warp owns 16 pixels (2 rows x 8 mel bins) x 16 channels per step with the
// lane layout and gate-GEMM fragments of bnglu_small.cu (lane (g, t4): mel bin g, channels 4 t4 .. 4 t4 + 3, both rows),
// and walks DOWN the time axis: a step reuses two of the four input rows under its pixels, all column predicates /
// addresses are loop invariants, the stencil and the accumulations run on packed fp32 (FFMA2: two channels per
// instruction), and with p = 0.5 one Philox call yields the 1-bit keep flags of 16 steps.  (The first version mapped
// groups to warps grid-stride like bnglu_small.cu: ncu counted 520 / 790 warp-instructions per group in forward /
// backward, over half of them address arithmetic, bounds tests and Philox - slower than the kernels it replaced.)
// Dropout counters depend only on (column of groups, lane, row block): forward and backward agree whatever their grids.
// The stencil keeps conv0_fwd's FMA order per channel, so z - and with it the eval-mode forward - is bit-identical to the
// unfused kernels (tests/test_layer0_gpu.py).
#include "bnglu_small.cuh"

namespace sedk {
namespace {

constexpr int L0C = 16;                 // filters of the first layer
constexpr int L0_GX = 0, L0_ZX = 144, L0_SX = 288;      // offsets (doubles) inside sedk_crnn_plan.l0_sums
constexpr int P_TT = 30;                // output rows per CTA of the x0 kernel (32-row halo: 128-byte loads along time)
constexpr int P_FW = 128;               // pixel columns per CTA
constexpr int P_HS = 129;               // tile row stride (odd: conflict-free transposing store)

// ---------------------------------------------------------------------------------------------------------------------
// sums != NULL: also the nine sums from which SX[tap] = sum over output pixels of x0(pix + tap) follows in closed form
// (l0_finish: the shifted window covers everything except one border row and / or column): total, first / last row,
// first / last column, four corners - accumulated over the batch in sums[L0_SX ..].
__global__ void __launch_bounds__(256)
l0_x0_kernel(const float* __restrict__ x, int64_t sb, int64_t sm, int64_t st, const uint32_t* __restrict__ minmax,
             float scaler_eps, const int32_t* __restrict__ specaug, float* __restrict__ x0, double* __restrict__ sums, int T,
             int F) {
    pdl_enter();
    __shared__ float til[32 * P_HS];
    __shared__ float s_b[9];
    const int tid = threadIdx.x;
    const int nTf = (F + P_FW - 1) / P_FW;
    const int nTt = (T + 31) / 32;
    int tile = blockIdx.x;
    const int b = tile / (nTt * nTf);
    tile -= b * nTt * nTf;
    const int t0 = (tile / nTf) * 32, f0 = (tile % nTf) * P_FW;
    float mn = 0.f, den = 1.f;
    const bool scale = minmax != nullptr;
    if (scale) {
        mn = ord2f(minmax[2 * b]);
        den = ord2f(minmax[2 * b + 1]) - mn + scaler_eps;
    }
    int fs = 0, fe = 0, ts = 0, te = 0;
    if (specaug) {
        fs = specaug[4 * b]; fe = specaug[4 * b + 1]; ts = specaug[4 * b + 2]; te = specaug[4 * b + 3];
    }
    const float* xb = x + (size_t)b * sb;
#pragma unroll 8
    for (int idx = tid; idx < 32 * P_FW; idx += 256) {
        int hr, hc;
        if (st == 1) { hc = idx >> 5; hr = idx & 31; }           // lanes run along time (reference layout [B, mel, T])
        else         { hr = idx / P_FW; hc = idx - hr * P_FW; }  // lanes run along mel (time-major input)
        const int t = t0 + hr, f = f0 + hc;
        float y = 0.f;
        if (t < T && f < F) {
            y = xb[(int64_t)f * sm + (int64_t)t * st];
            if (scale) y = (y - mn) / den * 2.0f - 1.0f;        // same operation order as TorchScaler (scaler.py:114-120)
            if ((f >= fs && f < fe) || (t >= ts && t < te)) y = 0.f;
        }
        til[hr * P_HS + hc] = y;
    }
    if (tid < 9) s_b[tid] = 0.f;
    __syncthreads();
    float tot = 0.f, r0 = 0.f, rl = 0.f, c0 = 0.f, cl = 0.f;
    for (int idx = tid; idx < 32 * P_FW; idx += 256) {
        const int r = idx / P_FW, c = idx - r * P_FW;
        const int t = t0 + r, f = f0 + c;
        if (t < T && f < F) {
            const float y = til[r * P_HS + c];
            x0[((size_t)b * T + t) * F + f] = y;
            tot += y;
            if (t == 0) r0 += y;
            if (t == T - 1) rl += y;
            if (f == 0) c0 += y;
            if (f == F - 1) cl += y;
            if (sums != nullptr && (t == 0 || t == T - 1) && (f == 0 || f == F - 1))
                atomicAdd(&s_b[5 + (t == 0 ? 0 : 2) + (f == 0 ? 0 : 1)], y);
        }
    }
    if (sums == nullptr) return;
    tot = warp_sum(tot); r0 = warp_sum(r0); rl = warp_sum(rl); c0 = warp_sum(c0); cl = warp_sum(cl);
    if ((tid & 31) == 0) {
        atomicAdd(&s_b[0], tot); atomicAdd(&s_b[1], r0); atomicAdd(&s_b[2], rl); atomicAdd(&s_b[3], c0); atomicAdd(&s_b[4], cl);
    }
    __syncthreads();
    if (tid < 9 && s_b[tid] != 0.f) atomicAdd(&sums[L0_SX + tid], (double)s_b[tid]);
}

// ---------------------------------------------------------------------------------------------------------------------
// strip decomposition: cols = B * F / 8 columns of groups, each cut into cpc chunks of L steps (1 step = 2 input rows)
struct Strips {
    int T, F, To, gpr;     // gpr = F / 8 groups per row
    int cpc, L, items;     // chunks per column, steps per chunk, cols * cpc
};

struct Walk {
    int col, b, fg, trow, trow1;
    const float* prow;     // x0 + (b T + 0) F + (8 fg - 1 + min(lane, 9)), clamped into the row: lanes 0..9 fetch the ten
    float mcol;            // columns under the warp's 8 pixels (+ halo), one load per row; 0 for a column outside the image
};

__device__ __forceinline__ Walk walk_begin(const Strips& sp, int item, const float* __restrict__ x0, int lane) {
    Walk w;
    w.col = item / sp.cpc;
    const int ch = item - w.col * sp.cpc;
    w.b = w.col / sp.gpr;
    w.fg = w.col - w.b * sp.gpr;
    w.trow = ch * sp.L;
    w.trow1 = min(sp.To, w.trow + sp.L);
    const int f = 8 * w.fg - 1 + min(lane, 9);
    w.prow = x0 + (size_t)w.b * sp.T * sp.F + min(max(f, 0), sp.F - 1);
    w.mcol = (f >= 0 && f < sp.F) ? 1.f : 0.f;
    return w;
}

typedef float2 Row[3];     // one input row under the lane (3 columns), each value in both halves of a packed pair

// A row is fetched one step ahead as a raw value (fetch_row: ONE load per row for the whole warp + the mask, nothing that
// waits for the data) and turned into the lane's three masked packed pairs only after the current step's arithmetic
// (finish_row: the mask multiply and three shuffles from lanes g, g + 1, g + 2): the first version loaded 3 values per
// lane and multiplied by the masks right behind the loads, stalling every step for the full L2 latency.
// Zero outside the image, branch-free: clamped addresses, 0 / 1 masks.  LOW: t may be negative (first row of a column).
struct RawRow {
    float v, m;
};
template <bool LOW>
__device__ __forceinline__ void fetch_row(RawRow& q, const Walk& w, const Strips& sp, int t) {
    q.m = ((!LOW || t >= 0) && t < sp.T) ? w.mcol : 0.f;
    q.v = __ldg(w.prow + min(LOW ? max(t, 0) : t, sp.T - 1) * sp.F);       // offset < 2^31: make_strips bounds B T F
}
__device__ __forceinline__ void finish_row(Row& r, const RawRow& q, int g) {
    const float v = q.v * q.m;
    const float l = __shfl_sync(0xffffffffu, v, g), c = __shfl_sync(0xffffffffu, v, g + 1), rr = __shfl_sync(0xffffffffu, v, g + 2);
    r[0] = make_float2(l, l);
    r[1] = make_float2(c, c);
    r[2] = make_float2(rr, rr);
}

// Walk the item's steps with a ring of six rows: step(ra, rb, rc, rd, trow) sees the four rows 2 trow - 1 .. 2 trow + 2
// while the two rows of the next step are in flight; three inlined copies of the body rotate the ring without moves.
template <class Step>
__device__ __forceinline__ void walk_rows(const Walk& w, const Strips& sp, int g, Step&& step) {
    Row r0, r1, r2, r3, r4, r5;
    RawRow qa, qb;
    {
        RawRow q0, q1;
        fetch_row<true>(q0, w, sp, 2 * w.trow - 1);
        fetch_row<true>(q1, w, sp, 2 * w.trow);
        fetch_row<true>(qa, w, sp, 2 * w.trow + 1);
        fetch_row<true>(qb, w, sp, 2 * w.trow + 2);
        finish_row(r0, q0, g); finish_row(r1, q1, g); finish_row(r2, qa, g); finish_row(r3, qb, g);
    }
    int trow = w.trow;
    while (true) {
        fetch_row<false>(qa, w, sp, 2 * trow + 3); fetch_row<false>(qb, w, sp, 2 * trow + 4);
        step(r0, r1, r2, r3, trow);
        finish_row(r4, qa, g); finish_row(r5, qb, g);
        if (++trow >= w.trow1) break;
        fetch_row<false>(qa, w, sp, 2 * trow + 3); fetch_row<false>(qb, w, sp, 2 * trow + 4);
        step(r2, r3, r4, r5, trow);
        finish_row(r0, qa, g); finish_row(r1, qb, g);
        if (++trow >= w.trow1) break;
        fetch_row<false>(qa, w, sp, 2 * trow + 3); fetch_row<false>(qb, w, sp, 2 * trow + 4);
        step(r4, r5, r0, r1, trow);
        finish_row(r2, qa, g); finish_row(r3, qb, g);
        if (++trow >= w.trow1) break;
    }
}

struct L0Weights {
    float2 w2[2][9];       // [channel pair][tap] = {w[4 t4 + 2 ep][tap], w[4 t4 + 2 ep + 1][tap]}
    float2 b2[2];
    __device__ __forceinline__ void load(const float* __restrict__ w, const float* __restrict__ bias, int t4) {
#pragma unroll
        for (int ep = 0; ep < 2; ep++) {
            const int ch = 4 * t4 + 2 * ep;
            b2[ep] = make_float2(__ldg(bias + ch), __ldg(bias + ch + 1));
#pragma unroll
            for (int k = 0; k < 9; k++) w2[ep][k] = make_float2(__ldg(w + ch * 9 + k), __ldg(w + (ch + 1) * 9 + k));
        }
    }
};

// z of one pixel row x 4 channels (packed pairs) from the three input rows around it; conv0_fwd_kernel's FMA order
__device__ __forceinline__ void stencil_row(float2 (&z)[2], const Row& ra, const Row& rb, const Row& rc, const L0Weights& W) {
#pragma unroll
    for (int ep = 0; ep < 2; ep++) {
        float2 a = W.b2[ep];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            a = __ffma2_rn(W.w2[ep][k], ra[k], a);
            a = __ffma2_rn(W.w2[ep][3 + k], rb[k], a);
            a = __ffma2_rn(W.w2[ep][6 + k], rc[k], a);
        }
        z[ep] = a;
    }
}
// acc[ep][tap] += v[ep] * x(tap) for one pixel row (the ZX / GX sums)
__device__ __forceinline__ void tap_sums(float2 (&acc)[2][9], const float2 (&v)[2], const Row& ra, const Row& rb, const Row& rc) {
#pragma unroll
    for (int ep = 0; ep < 2; ep++)
#pragma unroll
        for (int k = 0; k < 3; k++) {
            acc[ep][k] = __ffma2_rn(v[ep], ra[k], acc[ep][k]);
            acc[ep][3 + k] = __ffma2_rn(v[ep], rb[k], acc[ep][3 + k]);
            acc[ep][6 + k] = __ffma2_rn(v[ep], rc[k], acc[ep][6 + k]);
        }
}

// Keep flags of a lane's 8 elements (bit 4 rr + e) for the step at row block `trow`.
//   NB = 1  (p = 0.5): one Philox call per 16 steps, 1 bit per element (keep iff the bit is set);
//   NB = 16 (any p)  : one call per step, 16-bit draws against thresh16 (bnglu_small.cu's test).
template <int NB>
struct DropBits {
    uint4 r;
    __device__ __forceinline__ void refill(const Philox& ph, const Walk& w, int lane, int trow, uint64_t dstream) {
        const uint64_t who = (uint64_t)w.col * 32ull + (uint64_t)lane;
        r = ph((who << 11) | (uint64_t)(NB == 1 ? (trow >> 4) : trow), dstream);
    }
    __device__ __forceinline__ void begin(const Philox& ph, const Walk& w, int lane, uint64_t dstream) {
        if (NB == 1) refill(ph, w, lane, w.trow, dstream);
    }
    __device__ __forceinline__ uint32_t bits(const Philox& ph, const Walk& w, int lane, int trow, uint64_t dstream,
                                             uint32_t thresh16) {
        if (NB == 1) {
            if ((trow & 15) == 0) refill(ph, w, lane, trow, dstream);
            const int s = trow & 15;
            const uint32_t word = (s & 8) ? ((s & 4) ? r.w : r.z) : ((s & 4) ? r.y : r.x);
            return (word >> (8 * (s & 3))) & 0xffu;
        }
        refill(ph, w, lane, trow, dstream);
        const uint32_t wv[4] = {r.x, r.y, r.z, r.w};
        uint32_t b = 0;
#pragma unroll
        for (int e = 0; e < 4; e++) {
            b |= ((wv[e] & 0xffffu) >= thresh16 ? 1u : 0u) << e;
            b |= ((wv[e] >> 16) >= thresh16 ? 1u : 0u) << (4 + e);
        }
        return b;
    }
};

// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 3)
l0_stats_kernel(const float* __restrict__ x0, const float* __restrict__ w, const float* __restrict__ bias,
                double* __restrict__ stats, double* __restrict__ sums, Strips sp) {
    pdl_enter();
    constexpr int C = L0C;
    __shared__ float red[2 * C + C * 9];                      // sum z, sum z^2, ZX
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    for (int i = tid; i < 2 * C + C * 9; i += 128) red[i] = 0.f;
    L0Weights W;
    W.load(w, bias, t4);
    const float2 zero2 = make_float2(0.f, 0.f);
    float2 ssum[2] = {zero2, zero2}, ssq[2] = {zero2, zero2}, zx[2][9];
#pragma unroll
    for (int ep = 0; ep < 2; ep++)
#pragma unroll
        for (int k = 0; k < 9; k++) zx[ep][k] = zero2;
    __syncthreads();

    const int nwarps = gridDim.x * 4;
    for (int item = blockIdx.x * 4 + warp; item < sp.items; item += nwarps) {
        const Walk wk = walk_begin(sp, item, x0, lane);
        walk_rows(wk, sp, g, [&](const Row& ra, const Row& rb, const Row& rc, const Row& rd, int) {
            float2 z[2][2];
            stencil_row(z[0], ra, rb, rc, W);
            stencil_row(z[1], rb, rc, rd, W);
#pragma unroll
            for (int rr = 0; rr < 2; rr++)
#pragma unroll
                for (int ep = 0; ep < 2; ep++) {
                    ssum[ep] = __fadd2_rn(ssum[ep], z[rr][ep]);
                    ssq[ep] = __ffma2_rn(z[rr][ep], z[rr][ep], ssq[ep]);
                }
            tap_sums(zx, z[0], ra, rb, rc);
            tap_sums(zx, z[1], rb, rc, rd);
        });
    }
    // reduce over the 8 mel bins g of the warp (lanes with equal t4), then the CTA (shared atomics), then fp64 atomics
    auto over_g = [](float v) {
#pragma unroll
        for (int o = 4; o <= 16; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    };
#pragma unroll
    for (int ep = 0; ep < 2; ep++) {
        ssum[ep].x = over_g(ssum[ep].x); ssum[ep].y = over_g(ssum[ep].y);
        ssq[ep].x = over_g(ssq[ep].x); ssq[ep].y = over_g(ssq[ep].y);
#pragma unroll
        for (int k = 0; k < 9; k++) { zx[ep][k].x = over_g(zx[ep][k].x); zx[ep][k].y = over_g(zx[ep][k].y); }
    }
    if (g == 0) {
#pragma unroll
        for (int ep = 0; ep < 2; ep++) {
            const int ch = 4 * t4 + 2 * ep;
            atomicAdd(&red[ch], ssum[ep].x); atomicAdd(&red[ch + 1], ssum[ep].y);
            atomicAdd(&red[C + ch], ssq[ep].x); atomicAdd(&red[C + ch + 1], ssq[ep].y);
#pragma unroll
            for (int k = 0; k < 9; k++) {
                atomicAdd(&red[2 * C + ch * 9 + k], zx[ep][k].x);
                atomicAdd(&red[2 * C + (ch + 1) * 9 + k], zx[ep][k].y);
            }
        }
    }
    __syncthreads();
    if (tid < 2 * C) atomicAdd(&stats[tid], (double)red[tid]);
    for (int i = tid; i < C * 9; i += 128) atomicAdd(&sums[L0_ZX + i], (double)red[2 * C + i]);
}

// ---------------------------------------------------------------------------------------------------------------------
template <bool X3, int NB>
__global__ void __launch_bounds__(256, 2)
l0_fwd_kernel(const float* __restrict__ x0, const float* __restrict__ w, const float* __restrict__ bias,
              const float* __restrict__ bn, const float* __restrict__ glu_w, const float* __restrict__ glu_b,
              float* __restrict__ out, Strips sp, uint32_t thresh16, float inv_keep, uint64_t seed,
              const uint64_t* __restrict__ seed_dev, uint64_t dstream) {
    pdl_enter();
    constexpr int C = L0C;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    L0Weights W;
    W.load(w, bias, t4);
    float2 sc[2], sh[2];
    float bg[4];
#pragma unroll
    for (int ep = 0; ep < 2; ep++) {
        const int ch = 4 * t4 + 2 * ep;
        sc[ep] = make_float2(__ldg(bn + ch), __ldg(bn + ch + 1));
        sh[ep] = make_float2(__ldg(bn + C + ch), __ldg(bn + C + ch + 1));
        bg[2 * ep] = __ldg(glu_b + ch);
        bg[2 * ep + 1] = __ldg(glu_b + ch + 1);
    }
    GateB<C, X3> B1;
    B1.template load<1>(glu_w, g, t4);
    const Philox ph(seed + (seed_dev ? *seed_dev : 0ull));
    const bool drop = thresh16 != 0u;
    const int Fo = sp.F >> 1;

    const int nwarps = gridDim.x * 8;
    for (int item = blockIdx.x * 8 + warp; item < sp.items; item += nwarps) {
        Walk wk = walk_begin(sp, item, x0, lane);
        DropBits<NB> db;
        if (drop) db.begin(ph, wk, lane, dstream);
        // pooled output pixel of this lane pair: [b, trow, 4 fg + g / 2], channels 4 t4 ..
        float* op = out + (((size_t)wk.b * sp.To + wk.trow) * Fo + 4 * wk.fg + (g >> 1)) * C + 4 * t4;
        walk_rows(wk, sp, g, [&](const Row& ra, const Row& rb, const Row& rc, const Row& rd, int trow) {
            float2 z[2][2];
            stencil_row(z[0], ra, rb, rc, W);
            stencil_row(z[1], rb, rc, rd, W);
            float y[2][1][4];
#pragma unroll
            for (int rr = 0; rr < 2; rr++)
#pragma unroll
                for (int ep = 0; ep < 2; ep++) {
                    const float2 v = __ffma2_rn(z[rr][ep], sc[ep], sh[ep]);
                    y[rr][0][2 * ep] = v.x;
                    y[rr][0][2 * ep + 1] = v.y;
                }
            float acc[2][4];
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
                acc[0][2 * rr] = bg[0]; acc[0][2 * rr + 1] = bg[1];
                acc[1][2 * rr] = bg[2]; acc[1][2 * rr + 1] = bg[3];
            }
            gate_gemm<C, X3, 1>(acc, y, B1, glu_w, g, t4);
            uint32_t kb = 0xffu;
            if (drop) kb = db.bits(ph, wk, lane, trow, dstream, thresh16);
            float s[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                float v2[2];
#pragma unroll
                for (int rr = 0; rr < 2; rr++) {
                    // no branch on `drop` here (kb = 0xff, inv_keep = 1 without dropout): per-element branches cut the body
                    // into blocks of two sigmoids each and exposed the MUFU latency four times per step
                    const float v = acc[e >> 1][2 * rr + (e & 1)] * lean_sigmoidf(y[rr][0][e]);
                    v2[rr] = ((kb >> (4 * rr + e)) & 1u) ? v * inv_keep : 0.f;
                }
                s[e] = v2[0] + v2[1];
                s[e] = (s[e] + __shfl_xor_sync(0xffffffffu, s[e], 4)) * 0.25f;
                // TF32 mode: the pooled activation is the next convolution's MMA operand (the tcgen05 unit truncates)
                if (!X3) s[e] = __uint_as_float((__float_as_uint(s[e]) + 0x1000u) & 0xffffe000u);   // = cvt.rna.tf32 on finite values
            }
            if ((g & 1) == 0) *reinterpret_cast<float4*>(op) = make_float4(s[0], s[1], s[2], s[3]);
            op += (size_t)Fo * C;
        });
    }
}

// ---------------------------------------------------------------------------------------------------------------------
template <bool X3, int NB>
__global__ void __launch_bounds__(256)
l0_bwd_kernel(const float* __restrict__ x0, const float* __restrict__ w, const float* __restrict__ bias,
              const float* __restrict__ bn, const float* __restrict__ glu_w, const float* __restrict__ glu_b,
              const float* __restrict__ gout, float* __restrict__ gglu_w, float* __restrict__ gglu_b,
              double* __restrict__ stats, double* __restrict__ sums, Strips sp, uint32_t thresh16, float inv_keep,
              uint64_t seed, const uint64_t* __restrict__ seed_dev, uint64_t dstream) {
    pdl_enter();
    constexpr int C = L0C;
    constexpr int S = C + 8;                                  // row stride of the staging patch (bank-conflict-free reads)
    constexpr int PATCH = 2 * 16 * S;                         // floats per warp: g_lin[16][S], y[16][S]
    __shared__ __align__(16) float stage[8 * PATCH];
    __shared__ float red[3 * C + C * 9];                      // sum g_y, sum g_y*zhat, sum g_lin, GX
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    for (int i = tid; i < 3 * C + C * 9; i += 256) red[i] = 0.f;
    L0Weights W;
    W.load(w, bias, t4);
    float2 sc[2], sh[2], mi[2], is[2];
    float bg[4];
#pragma unroll
    for (int ep = 0; ep < 2; ep++) {
        const int ch = 4 * t4 + 2 * ep;
        sc[ep] = make_float2(__ldg(bn + ch), __ldg(bn + ch + 1));
        sh[ep] = make_float2(__ldg(bn + C + ch), __ldg(bn + C + ch + 1));
        is[ep] = make_float2(__ldg(bn + 3 * C + ch), __ldg(bn + 3 * C + ch + 1));
        mi[ep] = make_float2(-__ldg(bn + 2 * C + ch) * is[ep].x, -__ldg(bn + 2 * C + ch + 1) * is[ep].y);
        bg[2 * ep] = __ldg(glu_b + ch);
        bg[2 * ep + 1] = __ldg(glu_b + ch + 1);
    }
    GateB<C, X3> B1, B2;
    B1.template load<1>(glu_w, g, t4);
    B2.template load<2>(glu_w, g, t4);
    const Philox ph(seed + (seed_dev ? *seed_dev : 0ull));
    const bool drop = thresh16 != 0u;
    const float pool_keep = 0.25f * inv_keep;
    const int Fo = sp.F >> 1;
    float* Gs = stage + warp * PATCH;
    float* Ys = Gs + 16 * S;
    float dacc[2][4];
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
        for (int c = 0; c < 4; c++) dacc[j][c] = 0.f;
    const float2 zero2 = make_float2(0.f, 0.f);
    float2 s1[2] = {zero2, zero2}, s2[2] = {zero2, zero2}, sg[2] = {zero2, zero2}, gx[2][9];
#pragma unroll
    for (int ep = 0; ep < 2; ep++)
#pragma unroll
        for (int k = 0; k < 9; k++) gx[ep][k] = zero2;
    __syncthreads();

    const int nwarps = gridDim.x * 8;
    for (int item = blockIdx.x * 8 + warp; item < sp.items; item += nwarps) {
        Walk wk = walk_begin(sp, item, x0, lane);
        DropBits<NB> db;
        if (drop) db.begin(ph, wk, lane, dstream);
        const float* gp = gout + (((size_t)wk.b * sp.To + wk.trow) * Fo + 4 * wk.fg + (g >> 1)) * C + 4 * t4;
        float4 gon = __ldg(reinterpret_cast<const float4*>(gp));
        walk_rows(wk, sp, g, [&](const Row& ra, const Row& rb, const Row& rc, const Row& rd, int trow) {
            const float4 go = gon;
            gp += (size_t)Fo * C;
            if (trow + 1 < wk.trow1) gon = __ldg(reinterpret_cast<const float4*>(gp));
            float2 z[2][2];
            stencil_row(z[0], ra, rb, rc, W);
            stencil_row(z[1], rb, rc, rd, W);
            float y[2][1][4];
#pragma unroll
            for (int rr = 0; rr < 2; rr++)
#pragma unroll
                for (int ep = 0; ep < 2; ep++) {
                    const float2 v = __ffma2_rn(z[rr][ep], sc[ep], sh[ep]);
                    y[rr][0][2 * ep] = v.x;
                    y[rr][0][2 * ep + 1] = v.y;
                }
            float acc[2][4];
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
                acc[0][2 * rr] = bg[0]; acc[0][2 * rr + 1] = bg[1];
                acc[1][2 * rr] = bg[2]; acc[1][2 * rr + 1] = bg[3];
            }
            gate_gemm<C, X3, 1>(acc, y, B1, glu_w, g, t4);          // lin
            float gl[2][1][4];
            uint32_t kb = 0xffu;
            if (drop) kb = db.bits(ph, wk, lane, trow, dstream, thresh16);
#pragma unroll
            for (int rr = 0; rr < 2; rr++)
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    float ga = f4get(go, e) * pool_keep;            // 1/4 (average pool) x 1/(1 - p)
                    ga = ((kb >> (4 * rr + e)) & 1u) ? ga : 0.f;     // kb = 0xff without dropout: no branch in the body
                    const float sgm = lean_sigmoidf(y[rr][0][e]);
                    float& a = acc[e >> 1][2 * rr + (e & 1)];
                    const float g_lin = ga * sgm;
                    gl[rr][0][e] = g_lin;
                    a *= fmaf(-g_lin, sgm, g_lin);                   // lin * ga * s * (1 - s): elementwise part of g_y; GEMM 2 adds on top
                }
#pragma unroll
            for (int rr = 0; rr < 2; rr++)
#pragma unroll
                for (int ep = 0; ep < 2; ep++)
                    sg[ep] = __fadd2_rn(sg[ep], make_float2(gl[rr][0][2 * ep], gl[rr][0][2 * ep + 1]));
            // stage g_lin and y (pixel-major) for the gate weight-gradient GEMM
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
                *reinterpret_cast<float4*>(Gs + (g + 8 * rr) * S + 4 * t4) =
                    make_float4(gl[rr][0][0], gl[rr][0][1], gl[rr][0][2], gl[rr][0][3]);
                *reinterpret_cast<float4*>(Ys + (g + 8 * rr) * S + 4 * t4) =
                    make_float4(y[rr][0][0], y[rr][0][1], y[rr][0][2], y[rr][0][3]);
            }
            gate_gemm<C, X3, 2>(acc, gl, B2, glu_w, g, t4);         // g_y
            float2 gy[2][2];
#pragma unroll
            for (int rr = 0; rr < 2; rr++)
#pragma unroll
                for (int ep = 0; ep < 2; ep++) {
                    gy[rr][ep] = make_float2(acc[ep][2 * rr], acc[ep][2 * rr + 1]);
                    const float2 zh = __ffma2_rn(z[rr][ep], is[ep], mi[ep]);
                    s1[ep] = __fadd2_rn(s1[ep], gy[rr][ep]);
                    s2[ep] = __ffma2_rn(gy[rr][ep], zh, s2[ep]);
                }
            tap_sums(gx, gy[0], ra, rb, rc);
            tap_sums(gx, gy[1], rb, rc, rd);
            __syncwarp();
            // dWg[n][k] += sum_pix g_lin[pix][n] * y[pix][k]   (A = g_lin^T, B = y; K = the 16 pixels of the group)
#pragma unroll
            for (int ks = 0; ks < 2; ks++) {
                uint32_t bh[2][2], bl[2][2];
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const float b0 = Ys[(8 * ks + t4) * S + 8 * j + g], b1 = Ys[(8 * ks + t4 + 4) * S + 8 * j + g];
                    bh[j][0] = X3 ? to_tf32(b0) : to_tf32_mma(b0);
                    bh[j][1] = X3 ? to_tf32(b1) : to_tf32_mma(b1);
                    if (X3) {
                        bl[j][0] = to_tf32(b0 - __uint_as_float(bh[j][0]));
                        bl[j][1] = to_tf32(b1 - __uint_as_float(bh[j][1]));
                    }
                }
                const float a[4] = {Gs[(8 * ks + t4) * S + g], Gs[(8 * ks + t4) * S + g + 8],
                                    Gs[(8 * ks + t4 + 4) * S + g], Gs[(8 * ks + t4 + 4) * S + g + 8]};
                uint32_t ah[4], al[4];
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    ah[c] = X3 ? to_tf32(a[c]) : to_tf32_mma(a[c]);
                    if (X3) al[c] = to_tf32(a[c] - __uint_as_float(ah[c]));
                }
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    if (X3) {
                        mma_tf32(dacc[j], al, bh[j]);
                        mma_tf32(dacc[j], ah, bl[j]);
                    }
                    mma_tf32(dacc[j], ah, bh[j]);
                }
            }
            __syncwarp();
        });
    }

    // ---- flush: reduce over the 8 mel bins g of the warp, then the CTA (shared atomics), then fp64 global atomics
    auto over_g = [](float v) {
#pragma unroll
        for (int o = 4; o <= 16; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    };
#pragma unroll
    for (int ep = 0; ep < 2; ep++) {
        s1[ep].x = over_g(s1[ep].x); s1[ep].y = over_g(s1[ep].y);
        s2[ep].x = over_g(s2[ep].x); s2[ep].y = over_g(s2[ep].y);
        sg[ep].x = over_g(sg[ep].x); sg[ep].y = over_g(sg[ep].y);
#pragma unroll
        for (int k = 0; k < 9; k++) { gx[ep][k].x = over_g(gx[ep][k].x); gx[ep][k].y = over_g(gx[ep][k].y); }
    }
    if (g == 0) {
#pragma unroll
        for (int ep = 0; ep < 2; ep++) {
            const int ch = 4 * t4 + 2 * ep;
            atomicAdd(&red[ch], s1[ep].x); atomicAdd(&red[ch + 1], s1[ep].y);
            atomicAdd(&red[C + ch], s2[ep].x); atomicAdd(&red[C + ch + 1], s2[ep].y);
            atomicAdd(&red[2 * C + ch], sg[ep].x); atomicAdd(&red[2 * C + ch + 1], sg[ep].y);
#pragma unroll
            for (int k = 0; k < 9; k++) {
                atomicAdd(&red[3 * C + ch * 9 + k], gx[ep][k].x);
                atomicAdd(&red[3 * C + (ch + 1) * 9 + k], gx[ep][k].y);
            }
        }
    }
    __syncthreads();                                             // every warp is done with its staging patch
    for (int i = tid; i < C * C; i += 256) stage[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int n = g + 8 * (c >> 1), k = 8 * j + 2 * t4 + (c & 1);
            atomicAdd(&stage[n * C + k], dacc[j][c]);
        }
    __syncthreads();
    for (int i = tid; i < C * C; i += 256) atomicAdd(&gglu_w[i], stage[i]);
    for (int i = tid; i < C; i += 256) {
        atomicAdd(&stats[2 * C + i], (double)red[i]);
        atomicAdd(&stats[3 * C + i], (double)red[C + i]);
        atomicAdd(&gglu_b[i], red[2 * C + i]);
    }
    for (int i = tid; i < C * 9; i += 256) atomicAdd(&sums[L0_GX + i], (double)red[3 * C + i]);
}

// one CTA: the closed-form tail of the backward (header of this file)
__global__ void l0_finish_kernel(const double* __restrict__ stats, const double* __restrict__ sums,
                                 const float* __restrict__ bn, double inv_count, int frozen, float* __restrict__ gw,
                                 float* __restrict__ gb, float* __restrict__ ggamma, float* __restrict__ gbeta) {
    pdl_enter();
    constexpr int C = L0C;
    const int i = threadIdx.x;
    if (i < C * 9) {
        const int c = i / 9, tap = i - 9 * c;
        const double scale = (double)bn[c], mean = (double)bn[2 * C + c], invstd = (double)bn[3 * C + c];
        const double S1 = stats[2 * C + c], S2 = stats[3 * C + c];
        const double GX = sums[L0_GX + i];
        double v;
        if (frozen) {
            v = scale * GX;
        } else {
            // the window of tap (dy, dx) misses the last row (dy = 0) / first row (dy = 2) and the last / first column
            const double* b = sums + L0_SX;          // total, row 0, row T-1, column 0, column F-1, corners 00 0L L0 LL
            const int dy = tap / 3, dx = tap - 3 * dy;
            double SX = b[0];
            if (dy == 0) SX -= b[2];
            if (dy == 2) SX -= b[1];
            if (dx == 0) SX -= b[4];
            if (dx == 2) SX -= b[3];
            if (dy == 0 && dx == 0) SX += b[8];
            if (dy == 0 && dx == 2) SX += b[7];
            if (dy == 2 && dx == 0) SX += b[6];
            if (dy == 2 && dx == 2) SX += b[5];
            const double ZX = sums[L0_ZX + i];
            v = scale * (GX - S1 * inv_count * SX - S2 * inv_count * invstd * (ZX - mean * SX));
        }
        gw[i] += (float)v;
    }
    if (i < C) {
        const double S1 = stats[2 * C + i], S2 = stats[3 * C + i];
        gbeta[i] = (float)S1;
        ggamma[i] = (float)S2;
        // the conv bias cancels inside a batch-statistics BatchNorm; through a frozen one it sees scale * sum g_y
        gb[i] = frozen ? bn[i] * (float)S1 : 0.f;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// chunks per column: the cut that finishes first on `nwarps` persistent warps (waves x (steps per chunk + set-up))
bool make_strips(Strips& sp, int B, int T, int F, int nwarps) {
    if (F % 8 != 0 || T < 2 || (T & 1) || B < 1) return false;      // an odd last row would be pooled away but still counts in BN
    sp.T = T; sp.F = F; sp.To = T / 2; sp.gpr = F / 8;
    const long long cols = (long long)B * sp.gpr;
    if (cols * sp.To > 0x3fffffffLL || sp.To >= (1 << 11) || (long long)B * T * F > 0x7fffffffLL) return false;     // item index in 31 bits; Philox row field 11 bits
    long long best = -1;
    for (int want = 1; want <= sp.To; want++) {
        const int L = cdiv(sp.To, want), cpc = cdiv(sp.To, L);
        const long long items = cols * cpc;
        const long long cost = ((items + nwarps - 1) / nwarps) * (L + 3);
        if (best < 0 || cost < best) {
            best = cost;
            sp.cpc = cpc; sp.L = L; sp.items = (int)items;
        }
        if (items > 8LL * nwarps) break;
    }
    return true;
}

template <class K>
int walker_grid(K kern, int threads = 256) {
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, 0);
    return num_sms() * (occ < 1 ? 1 : occ);
}

// NB = 1 draws when the keep test is exactly one fair bit
inline bool fair_bit(float p) { return p == 0.5f; }

}  // namespace

bool l0_fused_supports(int B, int T, int F, int cout, int pt, int pf) {
    Strips sp;
    return cout == L0C && pt == 2 && pf == 2 && get_option("l0_fused", 1) != 0 && make_strips(sp, B, T, F, 148 * 8);
}

int launch_l0_prep(const float* x, int64_t sb, int64_t sm, int64_t st, const uint32_t* minmax, float scaler_eps,
                   const int32_t* specaug, const float* w, const float* bias, float* x0, double* stats, double* sums,
                   int B, int T, int F, cudaStream_t s) {
    SEDK_REQUIRE(x0 != nullptr && (stats == nullptr || sums != nullptr), "l0_prep: x0 / sums workspace missing");
    {
        SEDK_PROF("l0_x0", s);
        const int grid = B * cdiv(T, 32) * cdiv(F, P_FW);
        SEDK_CUDA(pdl_launch(l0_x0_kernel, dim3(grid), dim3(256), (size_t)(0), s, x, sb, sm, st, minmax, scaler_eps, specaug,
                             x0, stats != nullptr ? sums : nullptr, T, F));
        SEDK_LAUNCH_CHECK("l0_x0_kernel");
    }
    if (stats != nullptr) {
        SEDK_PROF("l0_stats", s);
        static int grid = 0;
        if (grid == 0) grid = walker_grid(l0_stats_kernel, 128);
        Strips sp;
        SEDK_REQUIRE(make_strips(sp, B, T, F, grid * 4), "l0_stats: unsupported geometry");
        const int need = cdiv(sp.items, 4);
        SEDK_CUDA(pdl_launch(l0_stats_kernel, dim3(grid < need ? grid : need), dim3(128), (size_t)(0), s, x0, w, bias, stats,
                             sums, sp));
        SEDK_LAUNCH_CHECK("l0_stats_kernel");
    }
    return SEDK_OK;
}

#define L0_DISPATCH(KERNEL, ...)                                                                            \
    do {                                                                                                    \
        const bool fb = fair_bit(drop_p);                                                                   \
        if (precision) {                                                                                    \
            if (fb) { L0_LAUNCH((KERNEL<true, 1>), __VA_ARGS__); } else { L0_LAUNCH((KERNEL<true, 16>), __VA_ARGS__); }   \
        } else {                                                                                            \
            if (fb) { L0_LAUNCH((KERNEL<false, 1>), __VA_ARGS__); } else { L0_LAUNCH((KERNEL<false, 16>), __VA_ARGS__); } \
        }                                                                                                   \
    } while (0)
#define L0_LAUNCH(K, ...)                                                                                   \
    do {                                                                                                    \
        static int grid = 0;                                                                                \
        if (grid == 0) grid = walker_grid(K);                                                               \
        Strips sp;                                                                                          \
        SEDK_REQUIRE(make_strips(sp, B, T, F, grid * 8), "layer0: unsupported geometry");                   \
        const int need = cdiv(sp.items, 8);                                                                 \
        SEDK_CUDA(pdl_launch(K, dim3(grid < need ? grid : need), dim3(256), (size_t)(0), s, __VA_ARGS__));  \
    } while (0)

int launch_l0_fwd(const float* x0, const float* w, const float* bias, const float* bn, const float* glu_w,
                  const float* glu_b, float* out, int B, int T, int F, float drop_p, uint64_t seed,
                  const uint64_t* seed_dev, uint64_t drop_stream, int precision, cudaStream_t s) {
    SEDK_PROF("l0_fwd", s);
    const uint32_t th = drop_threshold16(drop_p);
    const float inv_keep = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
    L0_DISPATCH(l0_fwd_kernel, x0, w, bias, bn, glu_w, glu_b, out, sp, th, inv_keep, seed, seed_dev, drop_stream);
    SEDK_LAUNCH_CHECK("l0_fwd_kernel");
    return SEDK_OK;
}

int launch_l0_bwd(const float* x0, const float* w, const float* bias, const float* bn, const float* glu_w,
                  const float* glu_b, const float* gout, float* gglu_w, float* gglu_b, double* stats, double* sums,
                  float* gw, float* gb, float* ggamma, float* gbeta, int B, int T, int F, int frozen, float drop_p,
                  uint64_t seed, const uint64_t* seed_dev, uint64_t drop_stream, int precision, cudaStream_t s) {
    const uint32_t th = drop_threshold16(drop_p);
    const float inv_keep = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
    {
        SEDK_PROF("l0_bwd", s);
        L0_DISPATCH(l0_bwd_kernel, x0, w, bias, bn, glu_w, glu_b, gout, gglu_w, gglu_b, stats, sums, sp, th, inv_keep, seed,
                    seed_dev, drop_stream);
        SEDK_LAUNCH_CHECK("l0_bwd_kernel");
    }
    {
        SEDK_PROF("l0_finish", s);
        const double inv_count = 1.0 / ((double)B * T * F);
        SEDK_CUDA(pdl_launch(l0_finish_kernel, dim3(1), dim3(160), (size_t)(0), s, stats, sums, bn, inv_count, frozen, gw, gb,
                             ggamma, gbeta));
        SEDK_LAUNCH_CHECK("l0_finish_kernel");
    }
    return SEDK_OK;
}

}  // namespace sedk
