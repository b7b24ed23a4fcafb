// Classification heads with attention pooling (desed_task/nnet/CRNN.py:152-178), the SED losses of
// SEDTask4.training_step (recipes/dcase2023_task4_baseline/local/sed_trainer.py:309-342), stand-alone dropout
// (CRNN.py:103,304) and the embedding fusion front (adaptive_avg_pool1d + concat + dropout, CRNN.py:280-294).
// All exact fp32 on CUDA cores: the contractions are tiny (C <= 32 classes) and sit right before the sigmoid/softmax.
#include "kernels.h"

namespace sedk {
namespace {

constexpr int HC_MAX = 32;     // classes
constexpr int HT = 32;         // time steps per chunk

// grid (time chunks of HT steps, clips).  hsum[b][0][c] += sum_t s*a, hsum[b][1][c] += sum_t a (zeroed by the launcher)
__global__ void __launch_bounds__(256)
heads_fwd_kernel(const float* __restrict__ x, const float* __restrict__ dw, const float* __restrict__ db,
                 const float* __restrict__ sw, const float* __restrict__ sb, const uint8_t* __restrict__ cmask,
                 float* __restrict__ strong, float* __restrict__ hsum, float* __restrict__ sof, int T, int D, int C) {
    pdl_enter();
    extern __shared__ float smem[];
    const int DS = D + 1;
    float* Wd = smem;                    // [C][DS]
    float* Ws = Wd + C * DS;             // [C][DS]
    float* xs = Ws + C * DS;             // [HT][DS]
    float* lg = xs + HT * DS;            // [HT][2*HC_MAX]
    const int b = blockIdx.y, tid = threadIdx.x;
    const int tc = blockIdx.x * HT;
    const int nt = min(HT, T - tc);
    for (int i = tid; i < C * D; i += 256) {
        int c = i / D, k = i - c * D;
        Wd[c * DS + k] = dw[i];
        Ws[c * DS + k] = sw[i];
    }
    const float* xb = x + (size_t)b * T * D;
    for (int i = tid; i < nt * D; i += 256) {
        int r = i / D, k = i - r * D;
        xs[r * DS + k] = xb[(size_t)(tc + r) * D + k];
    }
    __syncthreads();
    const int tl = tid >> 3, cg8 = tid & 7;
    if (tl < nt) {
        float ad[4] = {0.f, 0.f, 0.f, 0.f}, as[4] = {0.f, 0.f, 0.f, 0.f};
        const float* xr = xs + tl * DS;
        for (int k = 0; k < D; k++) {
            const float xv = xr[k];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int c = cg8 + 8 * q;
                if (c < C) {
                    ad[q] = fmaf(xv, Wd[c * DS + k], ad[q]);
                    as[q] = fmaf(xv, Ws[c * DS + k], as[q]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int c = cg8 + 8 * q;
            if (c < C) {
                lg[tl * 2 * HC_MAX + c] = ad[q] + db[c];
                lg[tl * 2 * HC_MAX + HC_MAX + c] = as[q] + sb[c];
            }
        }
    }
    __syncthreads();
    if (tid < 32) {                                      // warp 0: one lane per time step
        const bool live = tid < nt;
        const int t = tc + tid;
        const float* l = lg + tid * 2 * HC_MAX;
        float mx = -INFINITY, ssum = 0.f;
        if (live) {
            for (int c = 0; c < C; c++) {
                float v = l[HC_MAX + c];
                if (cmask && !cmask[b * C + c]) v = -1e30f;
                mx = fmaxf(mx, v);
            }
            for (int c = 0; c < C; c++) {
                float v = l[HC_MAX + c];
                if (cmask && !cmask[b * C + c]) v = -1e30f;
                ssum += expf(v - mx);
            }
        }
        for (int c = 0; c < C; c++) {
            float sa = 0.f, a = 0.f;
            if (live) {
                const bool ok = !cmask || cmask[b * C + c];
                const float v = ok ? l[HC_MAX + c] : -1e30f;
                const float p = expf(v - mx) / ssum;
                a = fminf(fmaxf(p, 1e-7f), 1.0f);
                const float sg = sigmoidf_(l[c]);
                sof[((size_t)b * T + t) * C + c] = p;
                strong[((size_t)b * C + c) * T + t] = ok ? sg : 0.f;
                sa = sg * a;
            }
            sa = warp_sum(sa);
            a = warp_sum(a);
            if (tid == 0) {
                atomicAdd(&hsum[(b * 2 + 0) * C + c], sa);
                atomicAdd(&hsum[(b * 2 + 1) * C + c], a);
            }
        }
    }
}

__global__ void heads_weak_kernel(const float* __restrict__ hsum, const uint8_t* __restrict__ cmask,
                                  float* __restrict__ weak, int B, int C) {
    pdl_enter();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * C) return;
    int b = i / C, c = i - b * C;
    const bool ok = !cmask || cmask[i];
    weak[i] = ok ? hsum[(b * 2) * C + c] / hsum[(b * 2 + 1) * C + c] : 0.f;
}

// grid (time chunks, clips); gl = grads wrt the two logit sets of this chunk
__global__ void __launch_bounds__(256)
heads_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dw, const float* __restrict__ sw,
                 const uint8_t* __restrict__ cmask, const float* __restrict__ strong, const float* __restrict__ hsum,
                 const float* __restrict__ sof, const float* __restrict__ gstrong, const float* __restrict__ gweak,
                 float* __restrict__ gx, float* __restrict__ gdw, float* __restrict__ gdb, float* __restrict__ gsw,
                 float* __restrict__ gsb, int T, int D, int C) {
    pdl_enter();
    __shared__ float gl[HT][2 * HC_MAX];
    extern __shared__ float wst[];                       // [2 * HC_MAX][256]: this thread's weight-gradient sums, by class
    const int b = blockIdx.y, tid = threadIdx.x;
    const int tc = blockIdx.x * HT;
    const int nt = min(HT, T - tc);
    {
        // 8 lanes per time step, classes strided over the lanes: every load of the step is in flight at once (one lane per
        // step walked the classes serially: ~20 dependent global-latency round trips before the GEMM part could start)
        const int tl = tid >> 3, cl = tid & 7;
        const bool live = tl < nt;
        const int t = tc + tl;
        float p4[4], s4[4], ga4[4], gst4[4], den4[4], gw4[4];
        float S = 0.f;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int c = cl + 8 * q;
            p4[q] = s4[q] = ga4[q] = gst4[q] = gw4[q] = 0.f;
            den4[q] = 1.f;
            if (live && c < C) {
                const bool ok = !cmask || cmask[b * C + c];
                const float p = sof[((size_t)b * T + t) * C + c];
                const float sv = strong[((size_t)b * C + c) * T + t];
                const float gw = (ok && gweak) ? gweak[b * C + c] : 0.f;
                const float gst = (ok && gstrong) ? gstrong[((size_t)b * C + c) * T + t] : 0.f;
                const float den = hsum[(b * 2 + 1) * C + c];
                const float wk = hsum[(b * 2) * C + c] / den;
                const float ga = gw * (sv - wk) / den;
                p4[q] = p; s4[q] = sv; ga4[q] = ga; gst4[q] = gst; den4[q] = den; gw4[q] = gw;
                const float gp = (p >= 1e-7f && p <= 1.0f) ? ga : 0.f;
                S += gp * p;
            }
        }
#pragma unroll
        for (int o = 1; o <= 4; o <<= 1) S += __shfl_xor_sync(0xffffffffu, S, o);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int c = cl + 8 * q;
            if (live && c < C) {
                const float p = p4[q], sv = s4[q];
                const float a = fminf(fmaxf(p, 1e-7f), 1.0f);
                const float gp = (p >= 1e-7f && p <= 1.0f) ? ga4[q] : 0.f;
                const float gs = gst4[q] + gw4[q] * a / den4[q];
                gl[tl][c] = gs * sv * (1.0f - sv);
                gl[tl][HC_MAX + c] = p * (gp - S);
            }
        }
    }
    __syncthreads();
    const float* xb = x + ((size_t)b * T + tc) * D;
    float* gxb = gx + ((size_t)b * T + tc) * D;
    for (int k = tid; k < D; k += 256) {
        float wd[HC_MAX], ws[HC_MAX], ad[HC_MAX], as[HC_MAX];
#pragma unroll
        for (int c = 0; c < HC_MAX; c++) {
            wd[c] = c < C ? dw[c * D + k] : 0.f;
            ws[c] = c < C ? sw[c * D + k] : 0.f;
            ad[c] = as[c] = 0.f;
        }
        // the x column of this thread is fetched 8 rows at a time (8 independent loads in flight instead of one per step)
        for (int tb = 0; tb < nt; tb += 8) {
            float xv8[8];
#pragma unroll
            for (int u = 0; u < 8; u++) xv8[u] = tb + u < nt ? xb[(size_t)(tb + u) * D + k] : 0.f;
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int t = tb + u;
                if (t < nt) {
                    const float xv = xv8[u];
                    float acc = 0.f;
#pragma unroll
                    for (int c = 0; c < HC_MAX; c++) {
                        if (c < C) {
                            const float g1 = gl[t][c], g2 = gl[t][HC_MAX + c];
                            acc = fmaf(g1, wd[c], acc);
                            acc = fmaf(g2, ws[c], acc);
                            ad[c] = fmaf(g1, xv, ad[c]);
                            as[c] = fmaf(g2, xv, as[c]);
                        }
                    }
                    gxb[(size_t)t * D + k] = acc;
                }
            }
        }
        // Every CTA adds into the same 2 C D addresses: started at class 0 everywhere, ~120 CTAs would queue on one L2
        // atomic unit at a time (measured: the atomics, not the math, set this kernel's time).  Stage the per-thread sums in
        // shared memory and let each CTA walk the classes from its own offset.
#pragma unroll
        for (int c = 0; c < HC_MAX; c++) {
            if (c < C) {
                wst[c * 256 + tid] = ad[c];
                wst[(HC_MAX + c) * 256 + tid] = as[c];
            }
        }
        const int rot = (blockIdx.x + blockIdx.y * gridDim.x) % C;
        for (int i = 0; i < C; i++) {
            int c = rot + i;
            if (c >= C) c -= C;
            atomicAdd(&gdw[c * D + k], wst[c * 256 + tid]);
            atomicAdd(&gsw[c * D + k], wst[(HC_MAX + c) * 256 + tid]);
        }
    }
    for (int c = tid; c < 2 * C; c += 256) {
        float sacc = 0.f;
        const int col = c < C ? c : HC_MAX + (c - C);
        for (int t = 0; t < nt; t++) sacc += gl[t][col];
        if (c < C) atomicAdd(&gdb[c], sacc);
        else atomicAdd(&gsb[c - C], sacc);
    }
}

__global__ void __launch_bounds__(256)
dropout_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, uint32_t thresh, float inv_keep,
               uint64_t seed, const uint64_t* __restrict__ seed_dev, uint64_t stream_id) {
    pdl_enter();
    const Philox ph(seed + (seed_dev ? *seed_dev : 0ull));
    const int64_t n4 = (n + 3) / 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        uint4 r = ph((uint64_t)i, stream_id);
        const uint32_t rv[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int64_t e = i * 4 + k;
            if (e < n) y[e] = rv[k] >= thresh ? x[e] * inv_keep : 0.f;
        }
    }
}

// SED losses + gradients wrt the posteriors. sums: [0] bce strong, [1] bce weak, [2] mse strong, [3] mse weak,
// [4] teacher bce strong, [5] teacher bce weak
__device__ __forceinline__ float bce_term(float p, float y) {
    return -(y * fmaxf(logf(p), -100.f) + (1.f - y) * fmaxf(logf(1.f - p), -100.f));
}
__device__ __forceinline__ float bce_grad(float p, float y) { return (p - y) / fmaxf((1.f - p) * p, 1e-12f); }

__global__ void __launch_bounds__(256)
sed_loss_kernel(const float* __restrict__ strong, const float* __restrict__ weak, const float* __restrict__ ts,
                const float* __restrict__ tw, const float* __restrict__ labels, const float* __restrict__ lweak, int B,
                int C, int T, int n_strong, int n_weak, int cons_row0, int cons_bce, float cw,
                const float* __restrict__ cw_dev, float* __restrict__ sums, float* __restrict__ gstrong,
                float* __restrict__ gweak) {
    pdl_enter();
    if (cw_dev != nullptr) cw = *cw_dev;
    const int64_t ns = (int64_t)B * C * T, nw = (int64_t)B * C;
    const float inv_bs = n_strong > 0 ? 1.f / (float)((int64_t)n_strong * C * T) : 0.f;
    const float inv_bw = n_weak > 0 ? 1.f / (float)((int64_t)n_weak * C) : 0.f;
    // consistency term over rows [cons_row0, B): all rows in the 2023 recipe, `mask_unlabeled` in the 2024 one
    const float inv_ms = 1.f / (float)((int64_t)(B - cons_row0) * C * T), inv_mw = 1.f / (float)((int64_t)(B - cons_row0) * C);
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ns + nw; i += (int64_t)gridDim.x * blockDim.x) {
        if (i < ns) {
            const int b = (int)(i / ((int64_t)C * T));
            const float p = strong[i];
            float g = 0.f;
            if (b < n_strong) {
                const float y = labels[i];
                acc[0] += bce_term(p, y);
                g += bce_grad(p, y) * inv_bs;
                if (ts) acc[4] += bce_term(ts[i], y);
            }
            if (ts && b >= cons_row0) {
                if (cons_bce) {
                    acc[2] += bce_term(p, ts[i]);
                    g += cw * bce_grad(p, ts[i]) * inv_ms;
                } else {
                    const float d = p - ts[i];
                    acc[2] += d * d;
                    g += cw * 2.f * d * inv_ms;
                }
            }
            if (gstrong) gstrong[i] = g;
        } else {
            const int64_t j = i - ns;
            const int b = (int)(j / C), c = (int)(j - (int64_t)b * C);
            const float p = weak[j];
            float g = 0.f;
            if (b >= n_strong && b < n_strong + n_weak) {
                const float y = lweak[(int64_t)(b - n_strong) * C + c];
                acc[1] += bce_term(p, y);
                g += bce_grad(p, y) * inv_bw;
                if (tw) acc[5] += bce_term(tw[j], y);
            }
            if (tw && b >= cons_row0) {
                if (cons_bce) {
                    acc[3] += bce_term(p, tw[j]);
                    g += cw * bce_grad(p, tw[j]) * inv_mw;
                } else {
                    const float d = p - tw[j];
                    acc[3] += d * d;
                    g += cw * 2.f * d * inv_mw;
                }
            }
            if (gweak) gweak[j] = g;
        }
    }
#pragma unroll
    for (int k = 0; k < 6; k++) {
        float v = warp_sum(acc[k]);
        if ((threadIdx.x & 31) == 0 && v != 0.f) atomicAdd(&sums[k], v);
    }
}

__global__ void sed_loss_finalize(float* losses, const float* sums, int B, int C, int T, int n_strong, int n_weak,
                                  int cons_row0, float cw, const float* cw_dev) {
    pdl_enter();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (cw_dev != nullptr) cw = *cw_dev;
    const float bs = n_strong > 0 ? sums[0] / (float)((int64_t)n_strong * C * T) : 0.f;
    const float bw = n_weak > 0 ? sums[1] / (float)((int64_t)n_weak * C) : 0.f;
    const float ms = sums[2] / (float)((int64_t)(B - cons_row0) * C * T);
    const float mw = sums[3] / (float)((int64_t)(B - cons_row0) * C);
    losses[1] = bs;
    losses[2] = bw;
    losses[3] = ms;
    losses[4] = mw;
    losses[5] = n_strong > 0 ? sums[4] / (float)((int64_t)n_strong * C * T) : 0.f;
    losses[6] = n_weak > 0 ? sums[5] / (float)((int64_t)n_weak * C) : 0.f;
    losses[7] = cw;
    losses[0] = bs + bw + (ms + mw) * cw;
}

// cat[b,t,:] = dropout( [ x[b,t,:] (x-span masked) , mean_{tau in window(t)} emb[b,:,tau] (e-span masked) ] )
__global__ void __launch_bounds__(256)
emb_concat_kernel(const float* __restrict__ x, const float* __restrict__ emb, const int32_t* __restrict__ dropstep,
                  float* __restrict__ cat, int T, int nb, int E, int Te, int mode, uint32_t thresh, float inv_keep,
                  uint64_t seed, const uint64_t* __restrict__ seed_dev, uint64_t stream_id) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int W = nb + E;
    const Philox ph(seed + (seed_dev ? *seed_dev : 0ull));
    int xs = 0, xe = 0, es = 0, ee = 0;
    if (dropstep) {
        xs = dropstep[4 * b]; xe = dropstep[4 * b + 1]; es = dropstep[4 * b + 2]; ee = dropstep[4 * b + 3];
    }
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    const int t0 = blockIdx.x * 32, e0 = blockIdx.y * 32;
    if (e0 < E) {
        // embedding part: read with lanes along time, write with lanes along the embedding axis
        for (int r = ty; r < 32; r += 8) {
            const int e = e0 + r, t = t0 + tx;
            float v = 0.f;
            if (e < E && t < T) {
                const float* ep = emb + ((size_t)b * E + e) * Te;
                if (mode == 1) {
                    // aggregation_type="interpolate" (CRNN.py:270-278): F.interpolate(mode="nearest-exact") along time,
                    // src = min(floor((dst + 0.5) * (float)(Te / T)), Te - 1) with ATen's float scale
                    const float scale = (float)Te / (float)T;
                    int q = (int)floorf((float)(((double)t + 0.5) * (double)scale));
                    v = ep[q < Te - 1 ? q : Te - 1];
                } else {
                    // aggregation_type="pool1d" (CRNN.py:280-283): adaptive_avg_pool1d windows
                    const int st = (int)(((int64_t)t * Te) / T);
                    const int en = (int)((((int64_t)(t + 1)) * Te + T - 1) / T);
                    float s = 0.f;
                    for (int q = st; q < en; q++) s += ep[q];
                    v = s / (float)(en - st);
                }
                if (t >= es && t < ee) v = 0.f;
            }
            tile[r][tx] = v;
        }
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {
            const int t = t0 + r, e = e0 + tx;
            if (t < T && e < E) {
                float v = tile[tx][r];
                const uint64_t idx = ((uint64_t)b * T + t) * W + nb + e;
                if (thresh != 0u) v = dropout_keep(ph, idx, stream_id, thresh) ? v * inv_keep : 0.f;
                cat[idx] = v;
            }
        }
    } else {
        // x part: blockIdx.y beyond the embedding tiles covers the nb CNN channels
        const int c0 = (blockIdx.y - (E + 31) / 32) * 32;
        for (int r = ty; r < 32; r += 8) {
            const int t = t0 + r, c = c0 + tx;
            if (t < T && c < nb) {
                float v = x[((size_t)b * T + t) * nb + c];
                if (t >= xs && t < xe) v = 0.f;
                const uint64_t idx = ((uint64_t)b * T + t) * W + c;
                if (thresh != 0u) v = dropout_keep(ph, idx, stream_id, thresh) ? v * inv_keep : 0.f;
                cat[idx] = v;
            }
        }
    }
}

__global__ void __launch_bounds__(256)
emb_concat_bwd_kernel(const float* __restrict__ gcat, const int32_t* __restrict__ dropstep, float* __restrict__ gx,
                      int T, int nb, int E, uint32_t thresh, float inv_keep, uint64_t seed,
                      const uint64_t* __restrict__ seed_dev, uint64_t stream_id, int64_t total) {
    const int W = nb + E;
    const Philox ph(seed + (seed_dev ? *seed_dev : 0ull));
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % nb);
        const int64_t bt = i / nb;
        const int t = (int)(bt % T), b = (int)(bt / T);
        const uint64_t idx = (uint64_t)bt * W + c;
        float v = gcat[idx];
        if (thresh != 0u) v = dropout_keep(ph, idx, stream_id, thresh) ? v * inv_keep : 0.f;
        if (dropstep && t >= dropstep[4 * b] && t < dropstep[4 * b + 1]) v = 0.f;
        gx[i] = v;
    }
}

}  // namespace

int launch_heads_fwd(const float* x, const float* dw, const float* db, const float* sw, const float* sb,
                     const uint8_t* cmask, float* strong, float* weak, float* sof, float* hsum, int B, int T, int D,
                     int C, cudaStream_t s) {
    SEDK_PROF("heads_fwd", s);
    SEDK_REQUIRE(C >= 1 && C <= HC_MAX, "heads: nclass %d must be in [1, %d]", C, HC_MAX);
    SEDK_REQUIRE(hsum != nullptr, "heads: hsum workspace missing");
    size_t smem = (size_t)(2 * C * (D + 1) + HT * (D + 1) + HT * 2 * HC_MAX) * sizeof(float);
    SEDK_REQUIRE(smem <= 227 * 1024, "heads: feature width %d too large", D);
    static size_t configured = 0;
    if (smem > configured) {
        int rc = opt_in_smem(heads_fwd_kernel, smem);
        if (rc) return rc;
        configured = smem;
    }
    SEDK_CUDA(cudaMemsetAsync(hsum, 0, (size_t)B * 2 * C * sizeof(float), s));
    dim3 grid(cdiv(T, HT), B);
    SEDK_CUDA(pdl_launch(heads_fwd_kernel, dim3(grid), dim3(256), (size_t)(smem), s, x, dw, db, sw, sb, cmask, strong, hsum, sof, T, D, C));
    SEDK_LAUNCH_CHECK("heads_fwd_kernel");
    SEDK_CUDA(pdl_launch(heads_weak_kernel, dim3(cdiv(B * C, 128)), dim3(128), (size_t)(0), s, hsum, cmask, weak, B, C));
    SEDK_LAUNCH_CHECK("heads_weak_kernel");
    return SEDK_OK;
}

int launch_heads_bwd(const float* x, const float* dw, const float* sw, const uint8_t* cmask, const float* strong,
                     const float* hsum, const float* sof, const float* gstrong, const float* gweak, float* gx,
                     float* gdw, float* gdb, float* gsw, float* gsb, int B, int T, int D, int C, cudaStream_t s) {
    SEDK_PROF("heads_bwd", s);
    SEDK_REQUIRE(C >= 1 && C <= HC_MAX, "heads: nclass %d must be in [1, %d]", C, HC_MAX);
    dim3 grid(cdiv(T, HT), B);
    const size_t smem = (size_t)2 * HC_MAX * 256 * sizeof(float);
    static bool configured = false;
    if (!configured) {
        int rc = opt_in_smem(heads_bwd_kernel, smem);
        if (rc) return rc;
        configured = true;
    }
    SEDK_CUDA(pdl_launch(heads_bwd_kernel, dim3(grid), dim3(256), (size_t)(smem), s, x, dw, sw, cmask, strong, hsum, sof, gstrong, gweak, gx, gdw, gdb, gsw, gsb, T,
                                             D, C));
    SEDK_LAUNCH_CHECK("heads_bwd_kernel");
    return SEDK_OK;
}

int launch_dropout(const float* x, float* y, int64_t n, float p, uint64_t seed, const uint64_t* seed_dev,
                   uint64_t stream_id, cudaStream_t s) {
    SEDK_PROF("dropout", s);
    const uint32_t thresh = drop_threshold(p);
    const float inv_keep = 1.0f / (1.0f - p);
    int64_t blocks = ((n + 3) / 4 + 255) / 256;
    int64_t cap = (int64_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    SEDK_CUDA(pdl_launch(dropout_kernel, dim3((int)blocks), dim3(256), (size_t)(0), s, x, y, n, thresh, inv_keep, seed, seed_dev, stream_id));
    SEDK_LAUNCH_CHECK("dropout_kernel");
    return SEDK_OK;
}

int launch_emb_concat(const float* x, const float* emb, const int32_t* dropstep, float* cat, int B, int T, int nb,
                      int emb_dim, int emb_T, int mode, float p, uint64_t seed, const uint64_t* seed_dev, uint64_t stream_id,
                      cudaStream_t s) {
    SEDK_PROF("emb_concat", s);
    const uint32_t thresh = p > 0.f ? drop_threshold(p) : 0u;
    const float inv_keep = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
    dim3 grid(cdiv(T, 32), cdiv(emb_dim, 32) + cdiv(nb, 32), B);
    emb_concat_kernel<<<grid, 256, 0, s>>>(x, emb, dropstep, cat, T, nb, emb_dim, emb_T, mode, thresh, inv_keep, seed,
                                           seed_dev, stream_id);
    SEDK_LAUNCH_CHECK("emb_concat_kernel");
    return SEDK_OK;
}

int launch_emb_concat_bwd(const float* gcat, const int32_t* dropstep, float* gx, int B, int T, int nb, int emb_dim,
                          float p, uint64_t seed, const uint64_t* seed_dev, uint64_t stream_id, cudaStream_t s) {
    SEDK_PROF("emb_concat_bwd", s);
    const uint32_t thresh = p > 0.f ? drop_threshold(p) : 0u;
    const float inv_keep = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
    const int64_t total = (int64_t)B * T * nb;
    int64_t blocks = (total + 255) / 256;
    int64_t cap = (int64_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    emb_concat_bwd_kernel<<<(int)blocks, 256, 0, s>>>(gcat, dropstep, gx, T, nb, emb_dim, thresh, inv_keep, seed, seed_dev,
                                                     stream_id, total);
    SEDK_LAUNCH_CHECK("emb_concat_bwd_kernel");
    return SEDK_OK;
}

}  // namespace sedk

static int sed_loss_impl(const float* strong, const float* weak, const float* t_strong, const float* t_weak,
                         const float* labels, const float* labels_weak, int B, int C, int T, int n_strong, int n_weak,
                         int cons_row0, int cons_kind, float cons_weight, const float* cw_dev, float* losses,
                         float* gstrong, float* gweak, void* stream) {
    using namespace sedk;
    SEDK_REQUIRE(cons_row0 >= 0 && cons_row0 < B && (cons_kind == 0 || cons_kind == 1), "sedk_sed_loss: bad cons_row0 / cons_kind");
    SEDK_REQUIRE(strong && weak && losses && B > 0 && C > 0 && T > 0, "sedk_sed_loss: bad arguments");
    SEDK_REQUIRE(n_strong >= 0 && n_weak >= 0 && n_strong + n_weak <= B, "sedk_sed_loss: n_strong + n_weak > B");
    SEDK_REQUIRE(n_strong == 0 || labels, "sedk_sed_loss: labels missing");
    SEDK_REQUIRE(n_weak == 0 || labels_weak, "sedk_sed_loss: labels_weak missing");
    SEDK_REQUIRE((t_strong == nullptr) == (t_weak == nullptr), "sedk_sed_loss: give both teacher tensors or none");
    cudaStream_t s = (cudaStream_t)stream;
    SEDK_PROF("sed_loss", s);
    SEDK_CUDA(cudaMemsetAsync(losses, 0, 16 * sizeof(float), s));      // [0,8) results, [8,16) running sums
    const int64_t n = (int64_t)B * C * T + (int64_t)B * C;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 4 * num_sms()) blocks = 4 * num_sms();
    SEDK_CUDA(pdl_launch(sed_loss_kernel, dim3(blocks), dim3(256), (size_t)(0), s, strong, weak, t_strong, t_weak, labels, labels_weak, B, C, T, n_strong, n_weak,
                                           cons_row0, cons_kind, cons_weight, cw_dev, losses + 8, gstrong, gweak));
    SEDK_LAUNCH_CHECK("sed_loss_kernel");
    SEDK_CUDA(pdl_launch(sed_loss_finalize, dim3(1), dim3(32), (size_t)(0), s, losses, losses + 8, B, C, T, n_strong, n_weak, cons_row0, cons_weight, cw_dev));
    SEDK_LAUNCH_CHECK("sed_loss_finalize");
    return SEDK_OK;
}

extern "C" int sedk_sed_loss(const float* strong, const float* weak, const float* t_strong, const float* t_weak,
                             const float* labels, const float* labels_weak, int B, int C, int T, int n_strong,
                             int n_weak, float cons_weight, float* losses, float* gstrong, float* gweak, void* stream) {
    return sed_loss_impl(strong, weak, t_strong, t_weak, labels, labels_weak, B, C, T, n_strong, n_weak, 0, 0, cons_weight,
                         nullptr, losses, gstrong, gweak, stream);
}

extern "C" int sedk_sed_loss_ex(const float* strong, const float* weak, const float* t_strong, const float* t_weak,
                                const float* labels, const float* labels_weak, int B, int C, int T, int n_strong,
                                int n_weak, int cons_row0, int cons_kind, float cons_weight, const float* cons_weight_dev,
                                float* losses, float* gstrong, float* gweak, void* stream) {
    return sed_loss_impl(strong, weak, t_strong, t_weak, labels, labels_weak, B, C, T, n_strong, n_weak, cons_row0,
                         cons_kind, cons_weight, cons_weight_dev, losses, gstrong, gweak, stream);
}

extern "C" int sedk_sed_loss_dev(const float* strong, const float* weak, const float* t_strong, const float* t_weak,
                                 const float* labels, const float* labels_weak, int B, int C, int T, int n_strong,
                                 int n_weak, const float* cons_weight_dev, float* losses, float* gstrong, float* gweak,
                                 void* stream) {
    if (cons_weight_dev == nullptr) {
        sedk::set_error("sedk_sed_loss_dev: cons_weight_dev is null");
        return SEDK_ERR_INVALID;
    }
    return sed_loss_impl(strong, weak, t_strong, t_weak, labels, labels_weak, B, C, T, n_strong, n_weak, 0, 0, 0.f,
                         cons_weight_dev, losses, gstrong, gweak, stream);
}
