// 3x3 convolutions of the CRNN front stack (desed_task/nnet/CNN.py:66-72) on channels-last tensors.
//
//  * conv0_fwd        Cin = 1: instance-minmax scaler (utils/scaler.py:114-120) + SpecAugment mask (CRNN.py:207-219)
//                     fused into the tile load, 9-tap stencil on CUDA cores, BN statistics in the epilogue.
//  * conv3x3_kernel   Cin >= 16: im2col-free implicit GEMM. The (TT+2)x(TF+2)xCin input halo is staged ONCE in shared
//                     memory (cp.async, zero-filled borders) and reused by all 9 taps; weights stream through a
//                     double-buffered cp.async ring; tensor-core mma.sync TF32 (3xTF32 in the fp32-parity mode);
//                     epilogue adds bias, stores, and reduces per-channel sum / sum^2 for train-mode BatchNorm.
//                     With the flipped/transposed weight pack the same kernel is the data-gradient.
//  * conv_wgrad       dW[tap][co][ci] = sum_pix gz[pix][co] * x[pix+tap][ci] as a TN tensor-core GEMM over pixel tiles
//                     with the x halo in shared memory, register accumulation across a persistent tile loop.
#include "kernels.h"

namespace sedk {
namespace {

// ------------------------------------------------------------------------------------------------------------
// rnd: round to the nearest TF32 value while packing (TF32 mode): the tcgen05 unit TRUNCATES fp32 operands, so operands
// that are already representable lose nothing there
__global__ void pack_kernel(const float* __restrict__ w, float* __restrict__ wp, int cin, int cout, int rnd) {
    const int n = 9 * cin * cout;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * n) return;
    int r = i % n;
    int tap = r / (cout * cin);
    float v;
    if (i < n) {
        int co = (r / cin) % cout, ci = r % cin;
        v = w[((size_t)co * cin + ci) * 9 + tap];
    } else {
        int ci = (r / cout) % cin, co = r % cout;
        v = w[((size_t)co * cin + ci) * 9 + (8 - tap)];
    }
    wp[i] = rnd ? __uint_as_float(to_tf32(v)) : v;
}

// Paired-pixel packs for a 16-channel layer (cin_l -> cout_l with one side = 16): two horizontally adjacent pixels are
// viewed as one pixel with twice the channels, [.., F, C] == [.., F / 2, 2 C], so that the 3x3 convolution becomes a
// 2 cin_l -> 2 cout_l convolution whose operand rows are >= 128 bytes (what the tcgen05 kernel tiles).  With output pixel
// (P, ho), input pixel (P + dP, hi) the original horizontal tap is dx = 2 dP + hi - ho + 1; taps outside {0, 1, 2} are zero.
//   wp[0 .. 9*2co*2ci)  forward  [tap'][(ho, co)][(hi, ci)] = w[co][ci][dy][dx],             tap' = dy * 3 + (dP + 1)
//   wp[9*2co*2ci .. )   dgrad    [tap'][(hi, ci)][(ho, co)] = w[co][ci][2 - dy][2 - dx'],    dx' = 2 dP + ho - hi + 1
__global__ void pack_pair_kernel(const float* __restrict__ w, float* __restrict__ wp, int cin, int cout) {
    const int n = 9 * 4 * cin * cout;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * n) return;
    const int r = i % n;
    const int tap = r / (4 * cin * cout), dy = tap / 3, dP = tap % 3 - 1;
    int co, ci, ho, hi, dx;
    float v = 0.f;
    if (i < n) {
        const int row = (r / (2 * cin)) % (2 * cout), col = r % (2 * cin);       // row = (ho, co), col = (hi, ci)
        ho = row / cout; co = row % cout; hi = col / cin; ci = col % cin;
        dx = 2 * dP + hi - ho + 1;
        if (dx >= 0 && dx <= 2) v = w[((size_t)co * cin + ci) * 9 + dy * 3 + dx];
    } else {
        const int row = (r / (2 * cout)) % (2 * cin), col = r % (2 * cout);      // row = (hi, ci), col = (ho, co)
        hi = row / cin; ci = row % cin; ho = col / cout; co = col % cout;
        dx = 2 * dP + ho - hi + 1;
        if (dx >= 0 && dx <= 2) v = w[((size_t)co * cin + ci) * 9 + (2 - dy) * 3 + (2 - dx)];
    }
    wp[i] = __uint_as_float(to_tf32(v));
}

__global__ void unpack_kernel(const float* __restrict__ gwp, float* __restrict__ gw, int cin, int cout) {
    const int n = 9 * cin * cout;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int tap = i % 9, ci = (i / 9) % cin, co = i / (9 * cin);
    gw[i] = gwp[((size_t)tap * cout + co) * cin + ci];
}

// ------------------------------------------------------------------------------------------------------------
// conv0: Cin = 1
constexpr int C0_TT = 30;      // output rows per CTA (halo 32 rows -> 128 B coalesced loads along time)
constexpr int C0_HS = 131;     // halo row stride (odd -> conflict-free transposing store)

template <int COUT>
__global__ void __launch_bounds__(256)
conv0_fwd_kernel(const float* __restrict__ x, int64_t sb, int64_t sm, int64_t st, const uint32_t* __restrict__ minmax,
                 float scaler_eps, const int32_t* __restrict__ specaug, const float* __restrict__ w,
                 const float* __restrict__ bias, float* __restrict__ x0, float* __restrict__ z,
                 double* __restrict__ stats, int T, int F) {
    pdl_enter();
    constexpr int TPP = COUT / 8;            // threads per pixel (8 channels each)
    constexpr int FW = 256 / TPP;            // pixel columns per CTA
    __shared__ float hal[(C0_TT + 2) * C0_HS];
    __shared__ float s_stat[2 * COUT];
    const int tid = threadIdx.x;
    const int nTf = (F + FW - 1) / FW;
    const int nTt = (T + C0_TT - 1) / C0_TT;
    int tile = blockIdx.x;
    const int b = tile / (nTt * nTf);
    tile -= b * nTt * nTf;
    const int t0 = (tile / nTf) * C0_TT, f0 = (tile % nTf) * FW;

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
    if (tid < 2 * COUT) s_stat[tid] = 0.f;
    const float* xb = x + (size_t)b * sb;
    constexpr int HR = C0_TT + 2, HC = FW + 2;
    // halo load in batches of 8 independent global loads per thread (one load per loop trip exposed the full memory latency
    // ~17 times per CTA: ncu attributed 23 % of the kernel's stall samples to the first use of the loaded value)
    constexpr int NLD = (HR * HC + 255) / 256;
#pragma unroll 1
    for (int base = 0; base < NLD; base += 8) {
        float v[8];
        int dst[8];
        bool msk[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int idx = tid + (base + u) * 256;
            int hr, hc;
            if (st == 1) { hc = idx / HR; hr = idx - hc * HR; }      // lanes run along time (reference layout)
            else         { hr = idx / HC; hc = idx - hr * HC; }      // lanes run along mel (time-major layout)
            const int t = t0 + hr - 1, f = f0 + hc - 1;
            const bool inb = idx < HR * HC;
            const bool ok = inb && t >= 0 && t < T && f >= 0 && f < F;
            dst[u] = inb ? hr * C0_HS + hc : -1;
            msk[u] = !ok || (f >= fs && f < fe) || (t >= ts && t < te);
            v[u] = ok ? xb[(int64_t)f * sm + (int64_t)t * st] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            float y = v[u];
            if (scale) y = (y - mn) / den * 2.0f - 1.0f;        // same operation order as TorchScaler (scaler.py:114-120)
            if (msk[u]) y = 0.f;
            if (dst[u] >= 0) hal[dst[u]] = y;
        }
    }
    __syncthreads();

    const int col = tid / TPP, cg = tid % TPP;      // pixel column, channel group
    const int f = f0 + col;
    float wr[8][9], br[8];
#pragma unroll
    for (int c = 0; c < 8; c++) {
        br[c] = bias[cg * 8 + c];
#pragma unroll
        for (int k = 0; k < 9; k++) wr[c][k] = w[(cg * 8 + c) * 9 + k];
    }
    float ssum[8], ssq[8];
#pragma unroll
    for (int c = 0; c < 8; c++) ssum[c] = ssq[c] = 0.f;
    const bool colok = (col < FW) && (f < F);
    if (colok) {
        float w0[3], w1[3], w2[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            w0[k] = hal[0 * C0_HS + col + k];
            w1[k] = hal[1 * C0_HS + col + k];
        }
        for (int r = 0; r < C0_TT; r++) {
            const int t = t0 + r;
            if (t >= T) break;
#pragma unroll
            for (int k = 0; k < 3; k++) w2[k] = hal[(r + 2) * C0_HS + col + k];
            float acc[8];
#pragma unroll
            for (int c = 0; c < 8; c++) {
                float a = br[c];
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    a = fmaf(wr[c][k], w0[k], a);
                    a = fmaf(wr[c][3 + k], w1[k], a);
                    a = fmaf(wr[c][6 + k], w2[k], a);
                }
                acc[c] = a;
                ssum[c] += a;
                ssq[c] = fmaf(a, a, ssq[c]);
            }
            const size_t pix = ((size_t)b * T + t) * F + f;
            float4* zp = reinterpret_cast<float4*>(z + pix * COUT + cg * 8);
            zp[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
            zp[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
            if (x0 != nullptr && cg == 0) x0[pix] = w1[1];
#pragma unroll
            for (int k = 0; k < 3; k++) { w0[k] = w1[k]; w1[k] = w2[k]; }
        }
    }
    if (stats != nullptr) {
        // lanes with equal (lane % TPP) hold the same channel group
#pragma unroll
        for (int c = 0; c < 8; c++) {
#pragma unroll
            for (int o = 16; o >= TPP; o >>= 1) {
                ssum[c] += __shfl_xor_sync(0xffffffffu, ssum[c], o);
                ssq[c] += __shfl_xor_sync(0xffffffffu, ssq[c], o);
            }
        }
        if ((tid & 31) < TPP) {
#pragma unroll
            for (int c = 0; c < 8; c++) {
                atomicAdd(&s_stat[cg * 8 + c], ssum[c]);
                atomicAdd(&s_stat[COUT + cg * 8 + c], ssq[c]);
            }
        }
        __syncthreads();
        if (tid < 2 * COUT) atomicAdd(&stats[tid], (double)s_stat[tid]);
    }
}

// conv0 weight gradient: TN tensor-core GEMM, M = cout, N = 9 taps (padded to 16), K = pixels
constexpr int W0_TT = 4, W0_FW = 128;
template <int COUT, bool X3>
__global__ void __launch_bounds__(256)
conv0_wgrad_kernel(const float* __restrict__ x0, const float* __restrict__ gz, float* __restrict__ gw, int B, int T,
                   int F, int total_tiles) {
    pdl_enter();
    constexpr int MF = COUT / 16;
    constexpr int GS = COUT + 8;
    constexpr int HS = W0_FW + 3;
    extern __shared__ float smem[];
    float* gzs = smem;                                   // [512][GS]
    float* xh = gzs + W0_TT * W0_FW * GS;                // [(TT+2)][HS]
    float* red = xh + (W0_TT + 2) * HS;                  // [COUT][16]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int nTf = (F + W0_FW - 1) / W0_FW, nTt = (T + W0_TT - 1) / W0_TT;
    float acc[MF][2][4];
#pragma unroll
    for (int i = 0; i < MF; i++)
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int q = 0; q < 4; q++) acc[i][j][q] = 0.f;
    for (int i = tid; i < COUT * 16; i += 256) red[i] = 0.f;

    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int r = tile;
        const int b = r / (nTt * nTf);
        r -= b * nTt * nTf;
        const int t0 = (r / nTf) * W0_TT, f0 = (r % nTf) * W0_FW;
        // gz tile (zero outside the image)
        for (int idx = tid; idx < W0_TT * W0_FW * (COUT / 4); idx += 256) {
            int p = idx / (COUT / 4), q = idx - p * (COUT / 4);
            int ty = p / W0_FW, tx = p - ty * W0_FW;
            int t = t0 + ty, f = f0 + tx;
            bool ok = (t < T) && (f < F);
            const float* src = ok ? gz + (((size_t)b * T + t) * F + f) * COUT + q * 4 : gz;
            cp_async16(gzs + p * GS + q * 4, src, ok);
        }
        cp_async_commit();
        for (int idx = tid; idx < (W0_TT + 2) * (W0_FW + 2); idx += 256) {
            int hr = idx / (W0_FW + 2), hc = idx - hr * (W0_FW + 2);
            int t = t0 + hr - 1, f = f0 + hc - 1;
            float v = 0.f;
            if (t >= 0 && t < T && f >= 0 && f < F) v = x0[((size_t)b * T + t) * F + f];
            xh[hr * HS + hc] = v;
        }
        cp_async_wait<0>();
        __syncthreads();
        // each warp: 64 pixels = 8 k8-steps
#pragma unroll 2
        for (int k8 = 0; k8 < 8; k8++) {
            const int pbase = warp * 64 + k8 * 8;
            auto fa = [&](int i, int r2, int c) {
                return gzs[(pbase + t4 + 4 * c) * GS + i * 16 + g + 8 * r2];
            };
            auto fb = [&](int j, int c) {
                const int p = pbase + t4 + 4 * c;
                const int ty = p / W0_FW, tx = p - ty * W0_FW;
                const int tap = j * 8 + g;
                if (tap > 8) return 0.f;
                const int dy = tap / 3, dx = tap - dy * 3;
                return xh[(ty + dy) * HS + tx + dx];
            };
            warp_mma_k8<MF, 2, X3>(acc, fa, fb);
        }
        __syncthreads();
    }
    // reduce the 8 warps in shared memory, then one atomic per output
#pragma unroll
    for (int i = 0; i < MF; i++)
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int q = 0; q < 4; q++) {
                int m = i * 16 + g + 8 * (q >> 1), n = j * 8 + 2 * t4 + (q & 1);
                atomicAdd(&red[m * 16 + n], acc[i][j][q]);
            }
    __syncthreads();
    for (int i = tid; i < COUT * 9; i += 256) {
        int m = i / 9, n = i - m * 9;
        atomicAdd(&gw[i], red[m * 16 + n]);
    }
}

// ------------------------------------------------------------------------------------------------------------
// generic 3x3 conv, Cin >= 16, tensor cores
template <int CIN, int NT, int TT, int TF>
struct ConvCfg {
    static constexpr int KC = CIN < 32 ? CIN : 32;
    static constexpr int HS = CIN + 4;
    static constexpr int BS = KC + 4;
    static constexpr int HW = TF + 2, HH = TT + 2, HP = HW * HH;
    static constexpr int WM = NT >= 64 ? 4 : 8, WN = 8 / WM;
    static constexpr int MF = (128 / WM) / 16, NF = (NT / WN) / 8;
    static constexpr size_t SMEM = (size_t)(HP * HS + 2 * NT * BS + 2 * NT) * sizeof(float);
    static_assert(TT * TF == 128, "tile must hold 128 pixels");
};

template <int CIN, int NT, int TT, int TF, bool X3>
__global__ void __launch_bounds__(256, 1)
conv3x3_kernel(const float* __restrict__ in, const float* __restrict__ wp, const float* __restrict__ bias,
               float* __restrict__ out, double* __restrict__ stats, int T, int F, int COUT) {
    pdl_enter();
    using Cfg = ConvCfg<CIN, NT, TT, TF>;
    constexpr int KC = Cfg::KC, HS = Cfg::HS, BS = Cfg::BS, HW = Cfg::HW, HP = Cfg::HP;
    constexpr int WN = Cfg::WN, MF = Cfg::MF, NF = Cfg::NF;
    extern __shared__ float smem[];
    float* halo = smem;
    float* Bs = halo + HP * HS;
    float* s_stat = Bs + 2 * NT * BS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int wm = warp / WN, wn = warp % WN;
    const int wm0 = wm * (MF * 16), wn0 = wn * (NF * 8);
    const int nTf = (F + TF - 1) / TF, nTt = (T + TT - 1) / TT;
    int tile = blockIdx.x;
    const int b = tile / (nTt * nTf);
    tile -= b * nTt * nTf;
    const int t0 = (tile / nTf) * TT, f0 = (tile % nTf) * TF;
    const int n0 = blockIdx.y * NT;

    if (tid < 2 * NT) s_stat[tid] = 0.f;
    // ---- halo (group 0)
    for (int idx = tid; idx < HP * (CIN / 4); idx += 256) {
        int p = idx / (CIN / 4), q = idx - p * (CIN / 4);
        int hy = p / HW, hx = p - hy * HW;
        int t = t0 + hy - 1, f = f0 + hx - 1;
        bool ok = (t >= 0) && (t < T) && (f >= 0) && (f < F);
        const float* src = ok ? in + (((size_t)b * T + t) * F + f) * CIN + q * 4 : in;
        cp_async16(halo + p * HS + q * 4, src, ok);
    }
    constexpr int NCH = CIN / KC, NIT = 9 * NCH;
    auto load_B = [&](int it, int buf) {
        const int tap = it / NCH, c0 = (it - tap * NCH) * KC;
        const float* src = wp + ((size_t)tap * COUT + n0) * CIN + c0;
        float* dst = Bs + buf * NT * BS;
        for (int idx = tid; idx < NT * (KC / 4); idx += 256) {
            int n = idx / (KC / 4), q = idx - n * (KC / 4);
            cp_async16(dst + n * BS + q * 4, src + (size_t)n * CIN + q * 4, true);
        }
    };
    load_B(0, 0);
    cp_async_commit();

    // ldmatrix row bases of this lane: A row (lane&7) + 8*((lane>>3)&1) of each m-fragment, k-offset 4*(lane>>4);
    // B row (lane&7) of fragment 2jp + (lane>>4), k-offset 4*((lane>>3)&1)
    uint32_t a_base[MF];
#pragma unroll
    for (int i = 0; i < MF; i++) {
        int m = wm0 + i * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
        int ty = m / TF, tx = m - ty * TF;
        a_base[i] = smem_u32(halo) + 4u * (uint32_t)((ty * HW + tx) * HS + 4 * (lane >> 4));
    }
    const uint32_t b_lane = 4u * (uint32_t)((wn0 + (lane >> 4) * 8 + (lane & 7)) * BS + 4 * ((lane >> 3) & 1));
    float acc[MF][NF][4];
#pragma unroll
    for (int i = 0; i < MF; i++)
#pragma unroll
        for (int j = 0; j < NF; j++)
#pragma unroll
            for (int q = 0; q < 4; q++) acc[i][j][q] = 0.f;

    for (int it = 0; it < NIT; it++) {
        if (it + 1 < NIT) load_B(it + 1, (it + 1) & 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const int tap = it / NCH, c0 = (it - tap * NCH) * KC;
        const int dy = tap / 3, dx = tap - dy * 3;
        const int toff = (dy * HW + dx) * HS + c0;
        const uint32_t bb = smem_u32(Bs + (it & 1) * NT * BS) + b_lane;
#pragma unroll
        for (int k8 = 0; k8 < KC / 8; k8++) {
            auto fa = [&](int i) { return a_base[i] + 4u * (uint32_t)(toff + k8 * 8); };
            auto fb = [&](int jp) { return bb + 4u * (uint32_t)(jp * 16 * BS + k8 * 8); };
            warp_mma_k8_ldsm<MF, NF, X3>(acc, fa, fb);
        }
        __syncthreads();
    }

    // ---- epilogue: bias, store, BN statistics
    float csum[NF][2], csq[NF][2];
#pragma unroll
    for (int j = 0; j < NF; j++) csum[j][0] = csum[j][1] = csq[j][0] = csq[j][1] = 0.f;
#pragma unroll
    for (int i = 0; i < MF; i++)
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const int m = wm0 + i * 16 + g + 8 * r;
            const int ty = m / TF, tx = m - ty * TF;
            const int t = t0 + ty, f = f0 + tx;
            if (t < T && f < F) {
                float* op = out + (((size_t)b * T + t) * F + f) * COUT + n0;
#pragma unroll
                for (int j = 0; j < NF; j++) {
                    const int n = wn0 + j * 8 + 2 * t4;
                    float v0 = acc[i][j][2 * r], v1 = acc[i][j][2 * r + 1];
                    if (bias != nullptr) { v0 += bias[n0 + n]; v1 += bias[n0 + n + 1]; }
                    *reinterpret_cast<float2*>(op + n) = make_float2(v0, v1);
                    csum[j][0] += v0; csum[j][1] += v1;
                    csq[j][0] = fmaf(v0, v0, csq[j][0]); csq[j][1] = fmaf(v1, v1, csq[j][1]);
                }
            }
        }
    if (stats != nullptr) {
#pragma unroll
        for (int j = 0; j < NF; j++)
#pragma unroll
            for (int q = 0; q < 2; q++) {
#pragma unroll
                for (int o = 4; o <= 16; o <<= 1) {
                    csum[j][q] += __shfl_xor_sync(0xffffffffu, csum[j][q], o);
                    csq[j][q] += __shfl_xor_sync(0xffffffffu, csq[j][q], o);
                }
                if (g == 0) {
                    const int n = wn0 + j * 8 + 2 * t4 + q;
                    atomicAdd(&s_stat[n], csum[j][q]);
                    atomicAdd(&s_stat[NT + n], csq[j][q]);
                }
            }
        __syncthreads();
        if (tid < NT) atomicAdd(&stats[n0 + tid], (double)s_stat[tid]);
        else if (tid < 2 * NT) atomicAdd(&stats[COUT + n0 + (tid - NT)], (double)s_stat[tid]);
    }
}

// ------------------------------------------------------------------------------------------------------------
// weight gradient
template <int CIN, int COUT, int TT, int TF, int NTAPS, int WM, int WN, int WK>
struct WgCfg {
    static constexpr int AS = COUT + 8, XS = CIN + 8;
    static constexpr int HW = TF + 2, HH = TT + 2, HP = HW * HH;
    static constexpr int MF = COUT / WM / 16, NF = CIN / WN / 8;
    static constexpr int KW = 128 / WK;
    static constexpr size_t SMEM =
        (size_t)(128 * AS + HP * XS + (WK > 1 ? NTAPS * COUT * CIN : 0)) * sizeof(float);
    static_assert(WM * WN * WK == 8, "8 warps");
    static_assert(TT * TF == 128, "tile must hold 128 pixels");
};

template <int CIN, int COUT, int TT, int TF, int NTAPS, int WM, int WN, int WK, bool X3>
__global__ void __launch_bounds__(256, 1)
conv_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ gz, float* __restrict__ gwp, int B, int T,
                  int F, int total_tiles) {
    using Cfg = WgCfg<CIN, COUT, TT, TF, NTAPS, WM, WN, WK>;
    constexpr int AS = Cfg::AS, XS = Cfg::XS, HW = Cfg::HW, HP = Cfg::HP, MF = Cfg::MF, NF = Cfg::NF, KW = Cfg::KW;
    extern __shared__ float smem[];
    float* gzs = smem;                 // [128][AS]
    float* xh = gzs + 128 * AS;        // [HP][XS]
    float* red = xh + HP * XS;         // [NTAPS][COUT][CIN] when WK > 1
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int wk = warp / (WM * WN), wm = (warp / WN) % WM, wn = warp % WN;
    const int wm0 = wm * (MF * 16), wn0 = wn * (NF * 8);
    const int tap0 = blockIdx.y * NTAPS;
    const int nTf = (F + TF - 1) / TF, nTt = (T + TT - 1) / TT;

    float acc[NTAPS][MF][NF][4];
#pragma unroll
    for (int a = 0; a < NTAPS; a++)
#pragma unroll
        for (int i = 0; i < MF; i++)
#pragma unroll
            for (int j = 0; j < NF; j++)
#pragma unroll
                for (int q = 0; q < 4; q++) acc[a][i][j][q] = 0.f;
    if (WK > 1)
        for (int i = tid; i < NTAPS * COUT * CIN; i += 256) red[i] = 0.f;

    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int r = tile;
        const int b = r / (nTt * nTf);
        r -= b * nTt * nTf;
        const int t0 = (r / nTf) * TT, f0 = (r % nTf) * TF;
        for (int idx = tid; idx < 128 * (COUT / 4); idx += 256) {
            int p = idx / (COUT / 4), q = idx - p * (COUT / 4);
            int ty = p / TF, tx = p - ty * TF;
            int t = t0 + ty, f = f0 + tx;
            bool ok = (t < T) && (f < F);
            const float* src = ok ? gz + (((size_t)b * T + t) * F + f) * COUT + q * 4 : gz;
            cp_async16(gzs + p * AS + q * 4, src, ok);
        }
        for (int idx = tid; idx < HP * (CIN / 4); idx += 256) {
            int p = idx / (CIN / 4), q = idx - p * (CIN / 4);
            int hy = p / HW, hx = p - hy * HW;
            int t = t0 + hy - 1, f = f0 + hx - 1;
            bool ok = (t >= 0) && (t < T) && (f >= 0) && (f < F);
            const float* src = ok ? x + (((size_t)b * T + t) * F + f) * CIN + q * 4 : x;
            cp_async16(xh + p * XS + q * 4, src, ok);
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
#pragma unroll
        for (int a = 0; a < NTAPS; a++) {
            const int tap = tap0 + a;
            const int dy = tap / 3, dx = tap - dy * 3;
#pragma unroll 2
            for (int k8 = 0; k8 < KW / 8; k8++) {
                const int kb = wk * KW + k8 * 8;
                auto fa = [&](int i, int r2, int c) { return gzs[(kb + t4 + 4 * c) * AS + wm0 + i * 16 + g + 8 * r2]; };
                auto fb = [&](int j, int c) {
                    const int p = kb + t4 + 4 * c;
                    const int ty = p / TF, tx = p - ty * TF;
                    return xh[((ty + dy) * HW + tx + dx) * XS + wn0 + j * 8 + g];
                };
                warp_mma_k8<MF, NF, X3>(acc[a], fa, fb);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < NTAPS; a++)
#pragma unroll
        for (int i = 0; i < MF; i++)
#pragma unroll
            for (int j = 0; j < NF; j++)
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const int m = wm0 + i * 16 + g + 8 * r, n = wn0 + j * 8 + 2 * t4;
                    if (WK > 1) {
                        atomicAdd(&red[(a * COUT + m) * CIN + n], acc[a][i][j][2 * r]);
                        atomicAdd(&red[(a * COUT + m) * CIN + n + 1], acc[a][i][j][2 * r + 1]);
                    } else {
                        float2* dst = reinterpret_cast<float2*>(gwp + ((size_t)(tap0 + a) * COUT + m) * CIN + n);
                        atomicAdd(dst, make_float2(acc[a][i][j][2 * r], acc[a][i][j][2 * r + 1]));
                    }
                }
    if (WK > 1) {
        __syncthreads();
        for (int i = tid; i < NTAPS * COUT * CIN; i += 256) atomicAdd(&gwp[(size_t)tap0 * COUT * CIN + i], red[i]);
    }
}

template <class K>
int launch_with_smem(K kernel, dim3 grid, size_t smem, cudaStream_t s, const char* name, bool& configured) {
    if (!configured) {
        int rc = opt_in_smem(kernel, smem);
        if (rc != SEDK_OK) return rc;
        configured = true;
    }
    (void)grid; (void)s; (void)name;
    return SEDK_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------
// Off by default.  Measured on B200 (layer 1 of the 2023 CRNN, 24 clips): the paired tcgen05 convolution takes 0.189 ms
// forward / 0.102 ms data-gradient against 0.079 / 0.065 ms for the halo-staged mma.sync kernel - with only 32 (paired)
// input channels the per-tap TMA re-fetch of the activation tile (9 x) and the 1920-CTA BatchNorm-statistics atomics cost
// more than the tensor pipe gains.  Kept as a parity-tested option ("conv_pair" = 1).
bool conv_pair_mode(int cin_l, int cout_l, int F, int precision) {
    return precision == 0 && tc5_enabled() && get_option("conv_pair", 0) != 0 && (F % 2) == 0 && F >= 4 &&
           ((cin_l == 16 && cout_l == 32) || (cin_l == 32 && cout_l == 16)) && tc5_supports(2 * cin_l, 2 * cout_l);
}
int conv_wpack_floats(int cin_l, int cout_l) {
    const int plain = 2 * 9 * cin_l * cout_l;
    return (cin_l == 16 || cout_l == 16) ? 4 * plain : plain;
}

int launch_pack_weights(const float* w, float* wpack, int cin, int cout, int round_tf32, cudaStream_t s) {
    SEDK_PROF("pack_weights", s);
    if (round_tf32 == 2) {          // paired-pixel packs (the caller asked conv_pair_mode)
        int n = 2 * 9 * 4 * cin * cout;
        pack_pair_kernel<<<cdiv(n, 256), 256, 0, s>>>(w, wpack, cin, cout);
        SEDK_LAUNCH_CHECK("pack_pair_kernel");
        return SEDK_OK;
    }
    int n = 2 * 9 * cin * cout;
    pack_kernel<<<cdiv(n, 256), 256, 0, s>>>(w, wpack, cin, cout, round_tf32);
    SEDK_LAUNCH_CHECK("pack_kernel");
    return SEDK_OK;
}
int launch_unpack_wgrad(const float* gwpack, float* gw, int cin, int cout, cudaStream_t s) {
    SEDK_PROF("unpack_wgrad", s);
    int n = 9 * cin * cout;
    unpack_kernel<<<cdiv(n, 256), 256, 0, s>>>(gwpack, gw, cin, cout);
    SEDK_LAUNCH_CHECK("unpack_kernel");
    return SEDK_OK;
}

int launch_conv0_fwd(const float* x, int64_t sb, int64_t sm, int64_t st, const uint32_t* minmax, float scaler_eps,
                     const int32_t* specaug, const float* w, const float* bias, float* x0, float* z, double* stats,
                     int B, int T, int F, int cout, cudaStream_t s) {
    SEDK_PROF("conv0_fwd", s);
    const int nTt = cdiv(T, C0_TT);
#define SEDK_C0(CO)                                                                                              \
    {                                                                                                            \
        const int FW = 256 / (CO / 8);                                                                           \
        dim3 grid(B * nTt * cdiv(F, FW));                                                                        \
        SEDK_CUDA(pdl_launch(conv0_fwd_kernel<CO>, dim3(grid), dim3(256), (size_t)(0), s, x, sb, sm, st, minmax, scaler_eps, specaug, w, bias, x0, z,    \
                                                  stats, T, F));                                                  \
    }
    if (cout == 16) SEDK_C0(16)
    else if (cout == 32) SEDK_C0(32)
    else if (cout == 64) SEDK_C0(64)
    else SEDK_UNSUPPORTED("conv0: first-layer width %d not in {16,32,64}", cout);
#undef SEDK_C0
    SEDK_LAUNCH_CHECK("conv0_fwd_kernel");
    return SEDK_OK;
}

int launch_conv0_wgrad(const float* x0, const float* gz, float* gw, int B, int T, int F, int cout, int precision,
                       cudaStream_t s) {
    SEDK_PROF("conv0_wgrad", s);
    const int tiles = B * cdiv(T, W0_TT) * cdiv(F, W0_FW);
    const int grid = tiles < 2 * num_sms() ? tiles : 2 * num_sms();
#define SEDK_W0(CO, X3)                                                                                          \
    {                                                                                                            \
        size_t smem = (size_t)(W0_TT * W0_FW * (CO + 8) + (W0_TT + 2) * (W0_FW + 3) + CO * 16) * sizeof(float);  \
        static bool cfg = false;                                                                                 \
        if (!cfg) { int rc = opt_in_smem(conv0_wgrad_kernel<CO, X3>, smem); if (rc) return rc; cfg = true; }    \
        SEDK_CUDA(pdl_launch(conv0_wgrad_kernel<CO, X3>, dim3(grid), dim3(256), (size_t)(smem), s, x0, gz, gw, B, T, F, tiles));                          \
    }
    if (cout == 16) { if (precision) SEDK_W0(16, true) else SEDK_W0(16, false) }
    else if (cout == 32) { if (precision) SEDK_W0(32, true) else SEDK_W0(32, false) }
    else if (cout == 64) { if (precision) SEDK_W0(64, true) else SEDK_W0(64, false) }
    else SEDK_UNSUPPORTED("conv0 wgrad: first-layer width %d not in {16,32,64}", cout);
#undef SEDK_W0
    SEDK_LAUNCH_CHECK("conv0_wgrad_kernel");
    return SEDK_OK;
}

template <int CIN, int NT, int TT, int TF>
static int run_conv(const float* in, const float* wp, const float* bias, float* out, double* stats, int B, int T,
                    int F, int cout, int precision, cudaStream_t s) {
    using Cfg = ConvCfg<CIN, NT, TT, TF>;
    dim3 grid(B * cdiv(T, TT) * cdiv(F, TF), cout / NT);
    static bool cfg0 = false, cfg1 = false;
    if (precision) {
        if (!cfg1) { int rc = opt_in_smem(conv3x3_kernel<CIN, NT, TT, TF, true>, Cfg::SMEM); if (rc) return rc; cfg1 = true; }
        SEDK_CUDA(pdl_launch(conv3x3_kernel<CIN, NT, TT, TF, true>, dim3(grid), dim3(256), (size_t)(Cfg::SMEM), s, in, wp, bias, out, stats, T, F, cout));
    } else {
        if (!cfg0) { int rc = opt_in_smem(conv3x3_kernel<CIN, NT, TT, TF, false>, Cfg::SMEM); if (rc) return rc; cfg0 = true; }
        SEDK_CUDA(pdl_launch(conv3x3_kernel<CIN, NT, TT, TF, false>, dim3(grid), dim3(256), (size_t)(Cfg::SMEM), s, in, wp, bias, out, stats, T, F, cout));
    }
    SEDK_LAUNCH_CHECK("conv3x3_kernel");
    return SEDK_OK;
}

template <int CIN, int NT>
static int run_conv_tiles(const float* in, const float* wp, const float* bias, float* out, double* stats, int B, int T,
                          int F, int cout, int precision, cudaStream_t s) {
    if (F > 8) return run_conv<CIN, NT, 8, 16>(in, wp, bias, out, stats, B, T, F, cout, precision, s);
    if (F > 4) return run_conv<CIN, NT, 16, 8>(in, wp, bias, out, stats, B, T, F, cout, precision, s);
    if (F > 2) return run_conv<CIN, NT, 32, 4>(in, wp, bias, out, stats, B, T, F, cout, precision, s);
    return run_conv<CIN, NT, 64, 2>(in, wp, bias, out, stats, B, T, F, cout, precision, s);
}

int launch_conv3x3_layer(const float* in, const float* wpack_base, int dgrad, const float* bias, float* out, double* stats,
                         int B, int T, int F, int cin_l, int cout_l, int precision, cudaStream_t s) {
    const int cin = dgrad ? cout_l : cin_l, cout = dgrad ? cin_l : cout_l;     // channels of THIS convolution
    if (conv_pair_mode(cin_l, cout_l, F, precision)) {
        const float* wp = wpack_base + (dgrad ? (size_t)9 * 4 * cin_l * cout_l : 0);
        return launch_conv3x3_tc5(in, wp, bias, out, stats, B, T, F / 2, 2 * cin, 2 * cout, cout, s);
    }
    return launch_conv3x3(in, wpack_base + (dgrad ? (size_t)9 * cin_l * cout_l : 0), bias, out, stats, B, T, F, cin, cout,
                          precision, s);
}

int launch_conv3x3(const float* in, const float* wp, const float* bias, float* out, double* stats, int B, int T, int F,
                   int cin, int cout, int precision, cudaStream_t s) {
    if (precision == 0 && tc5_enabled() && tc5_supports(cin, cout))
        return launch_conv3x3_tc5(in, wp, bias, out, stats, B, T, F, cin, cout, cout, s);
    char pname[64];
    snprintf(pname, sizeof(pname), "conv3x3_%dto%d_F%d", cin, cout, F);
    SEDK_PROF(pname, s);
    // 128-wide inputs: two 64-channel output tiles per pixel tile -> 113 KB of shared memory, 2 CTAs (16 warps) per SM
    const int nt = cout >= 128 ? (cin >= 128 ? 64 : 128) : cout;
    SEDK_REQUIRE(cout % nt == 0, "conv3x3: cout %d must be a multiple of %d", cout, nt);
#define SEDK_CONV(CI, NTV) \
    if (cin == CI && nt == NTV) return run_conv_tiles<CI, NTV>(in, wp, bias, out, stats, B, T, F, cout, precision, s);
    SEDK_CONV(16, 16) SEDK_CONV(16, 32) SEDK_CONV(32, 16) SEDK_CONV(32, 32) SEDK_CONV(32, 64)
    SEDK_CONV(64, 32) SEDK_CONV(64, 64) SEDK_CONV(64, 128) SEDK_CONV(128, 64) SEDK_CONV(128, 128)
#undef SEDK_CONV
    SEDK_UNSUPPORTED("conv3x3: (cin=%d, cout=%d) has no sm_100a instantiation (supported: channel widths 16/32/64/128 "
                     "with at most a 2x step between layers)", cin, cout);
}

template <int CIN, int COUT, int TT, int TF, int NTAPS, int WM, int WN, int WK>
static int run_wgrad(const float* x, const float* gz, float* gwp, int B, int T, int F, int precision, cudaStream_t s) {
    using Cfg = WgCfg<CIN, COUT, TT, TF, NTAPS, WM, WN, WK>;
    const int tiles = B * cdiv(T, TT) * cdiv(F, TF);
    const int groups = 9 / NTAPS;
    int gx = (2 * num_sms() + groups - 1) / groups;
    if (gx > tiles) gx = tiles;
    dim3 grid(gx, groups);
    static bool cfg0 = false, cfg1 = false;
    if (precision) {
        auto k = conv_wgrad_kernel<CIN, COUT, TT, TF, NTAPS, WM, WN, WK, true>;
        if (!cfg1) { int rc = opt_in_smem(k, Cfg::SMEM); if (rc) return rc; cfg1 = true; }
        k<<<grid, 256, Cfg::SMEM, s>>>(x, gz, gwp, B, T, F, tiles);
    } else {
        auto k = conv_wgrad_kernel<CIN, COUT, TT, TF, NTAPS, WM, WN, WK, false>;
        if (!cfg0) { int rc = opt_in_smem(k, Cfg::SMEM); if (rc) return rc; cfg0 = true; }
        k<<<grid, 256, Cfg::SMEM, s>>>(x, gz, gwp, B, T, F, tiles);
    }
    SEDK_LAUNCH_CHECK("conv_wgrad_kernel");
    return SEDK_OK;
}

template <int CIN, int COUT, int NTAPS, int WM, int WN, int WK>
static int run_wgrad_tiles(const float* x, const float* gz, float* gwp, int B, int T, int F, int precision,
                           cudaStream_t s) {
    if (F > 8) return run_wgrad<CIN, COUT, 8, 16, NTAPS, WM, WN, WK>(x, gz, gwp, B, T, F, precision, s);
    if (F > 4) return run_wgrad<CIN, COUT, 16, 8, NTAPS, WM, WN, WK>(x, gz, gwp, B, T, F, precision, s);
    if (F > 2) return run_wgrad<CIN, COUT, 32, 4, NTAPS, WM, WN, WK>(x, gz, gwp, B, T, F, precision, s);
    return run_wgrad<CIN, COUT, 64, 2, NTAPS, WM, WN, WK>(x, gz, gwp, B, T, F, precision, s);
}

int launch_conv_wgrad(const float* x, const float* gz, float* gwpack, int B, int T, int F, int cin, int cout,
                      int precision, cudaStream_t s) {
    if (precision == 0 && tc5_enabled() && tc5_wgrad_supports(cin, cout))
        return launch_conv_wgrad_tc5(x, gz, gwpack, B, T, F, cin, cout, s);
    char pname[64];
    snprintf(pname, sizeof(pname), "conv_wgrad_%dto%d_F%d", cin, cout, F);
    SEDK_PROF(pname, s);
    //                                         CIN  COUT NTAPS WM WN WK
    if (cin == 16 && cout == 32) return run_wgrad_tiles<16, 32, 9, 2, 1, 4>(x, gz, gwpack, B, T, F, precision, s);
    if (cin == 32 && cout == 64) return run_wgrad_tiles<32, 64, 9, 4, 2, 1>(x, gz, gwpack, B, T, F, precision, s);
    if (cin == 64 && cout == 128) return run_wgrad_tiles<64, 128, 3, 4, 2, 1>(x, gz, gwpack, B, T, F, precision, s);
    if (cin == 128 && cout == 128) return run_wgrad_tiles<128, 128, 1, 4, 2, 1>(x, gz, gwpack, B, T, F, precision, s);
    SEDK_UNSUPPORTED("conv wgrad: (cin=%d, cout=%d) has no sm_100a instantiation (supported: 16->32, 32->64, 64->128, "
                     "128->128)", cin, cout);
}

}  // namespace sedk

extern "C" int sedk_conv3x3(const float* in, const float* wpack, const float* bias, float* out, double* stats, int B, int T,
                            int F, int cin, int cout, int precision, void* stream) {
    using namespace sedk;
    SEDK_REQUIRE(in && wpack && out && B > 0 && T > 0 && F > 0, "sedk_conv3x3: bad arguments");
    return launch_conv3x3(in, wpack, bias, out, stats, B, T, F, cin, cout, precision, (cudaStream_t)stream);
}

extern "C" int sedk_conv_wgrad(const float* x, const float* gz, float* gwpack, int B, int T, int F, int cin, int cout,
                               int precision, void* stream) {
    using namespace sedk;
    SEDK_REQUIRE(x && gz && gwpack && B > 0 && T > 0 && F > 0, "sedk_conv_wgrad: bad arguments");
    return launch_conv_wgrad(x, gz, gwpack, B, T, F, cin, cout, precision, (cudaStream_t)stream);
}
