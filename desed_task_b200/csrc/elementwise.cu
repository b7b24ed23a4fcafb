// HBM-bound elementwise / reduction kernels: mixup + log + per-clip min/max, scaler modes, label mixup, frame shift,
// additive noise, fused Adam + EMA over a flat parameter buffer, median filter.
// All are streaming kernels: 128-bit loads/stores where alignment allows, grid sized in multiples of the SM count.
#include "common.cuh"
#include <cuda_bf16.h>

namespace sedk {
namespace {

constexpr int EW_THREADS = 256;

__device__ __forceinline__ float amp_to_db(float v, float amin, float lo, float hi) {
    float d = 20.0f * log10f(fmaxf(v, amin));
    return fminf(fmaxf(d, lo), hi);
}

__global__ void minmax_init_kernel(uint32_t* mm, int B) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) {
        mm[2 * i] = f2ord(INFINITY);
        mm[2 * i + 1] = f2ord(-INFINITY);
    }
}
__global__ void minmax_decode_kernel(const uint32_t* mm, float* out, int B) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 2 * B) out[i] = ord2f(mm[i]);
}

// grid = (chunks, B); each CTA streams a contiguous slice of one clip
template <bool VEC>
__global__ void __launch_bounds__(EW_THREADS)
feat_mix_log_kernel(const float* __restrict__ x, const int64_t* __restrict__ perm, const float* __restrict__ coef,
                    float* __restrict__ out, int64_t n, int log_mode, float amin, float lo, float hi,
                    uint32_t* __restrict__ minmax) {
    const int b = blockIdx.y;
    const float* xa = x + (size_t)b * n;
    const float* xb = perm ? x + (size_t)perm[b] * n : nullptr;
    const float c = (perm && coef) ? coef[b] : 1.0f;
    const float c1 = 1.0f - c;
    float* o = out + (size_t)b * n;
    float vmin = INFINITY, vmax = -INFINITY;
    if (VEC) {
        const int64_t n4 = n >> 2;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
            float4 a = reinterpret_cast<const float4*>(xa)[i];
            if (xb) {
                float4 p = reinterpret_cast<const float4*>(xb)[i];
                a.x = c * a.x + c1 * p.x; a.y = c * a.y + c1 * p.y; a.z = c * a.z + c1 * p.z; a.w = c * a.w + c1 * p.w;
            }
            if (log_mode) {
                a.x = amp_to_db(a.x, amin, lo, hi); a.y = amp_to_db(a.y, amin, lo, hi);
                a.z = amp_to_db(a.z, amin, lo, hi); a.w = amp_to_db(a.w, amin, lo, hi);
            }
            reinterpret_cast<float4*>(o)[i] = a;
            vmin = fminf(fminf(vmin, a.x), fminf(fminf(a.y, a.z), a.w));
            vmax = fmaxf(fmaxf(vmax, a.x), fmaxf(fmaxf(a.y, a.z), a.w));
        }
    } else {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
            float a = xa[i];
            if (xb) a = c * a + c1 * xb[i];
            if (log_mode) a = amp_to_db(a, amin, lo, hi);
            o[i] = a;
            vmin = fminf(vmin, a);
            vmax = fmaxf(vmax, a);
        }
    }
    if (minmax) {
        vmin = warp_min(vmin);
        vmax = warp_max(vmax);
        if ((threadIdx.x & 31) == 0) {
            atomicMin(minmax + 2 * b, f2ord(vmin));
            atomicMax(minmax + 2 * b + 1, f2ord(vmax));
        }
    }
}

__global__ void __launch_bounds__(EW_THREADS)
minmax_scale_kernel(const float* __restrict__ x, float* __restrict__ out, const uint32_t* __restrict__ minmax,
                    int64_t n, float eps) {
    const int b = blockIdx.y;
    const float mn = ord2f(minmax[2 * b]), mx = ord2f(minmax[2 * b + 1]);
    const float den = mx - mn + eps;
    const float* xa = x + (size_t)b * n;
    float* o = out + (size_t)b * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        o[i] = (xa[i] - mn) / den * 2.0f - 1.0f;
}

// one CTA per clip: two-pass mean / unbiased std in fp32 with fp64 block combine
__global__ void __launch_bounds__(EW_THREADS) instance_stats_kernel(const float* __restrict__ x, float* __restrict__ stats,
                                                                    int64_t n) {
    __shared__ double red[EW_THREADS / 32];
    __shared__ double s_mean;
    const int b = blockIdx.x;
    const float* xa = x + (size_t)b * n;
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) acc += (double)xa[i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int i = 0; i < EW_THREADS / 32; i++) t += red[i];
        s_mean = t / (double)n;
    }
    __syncthreads();
    const double mean = s_mean;
    double acc2 = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        double d = (double)xa[i] - mean;
        acc2 += d * d;
    }
    for (int o = 16; o > 0; o >>= 1) acc2 += __shfl_xor_sync(0xffffffffu, acc2, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc2;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int i = 0; i < EW_THREADS / 32; i++) t += red[i];
        stats[2 * b] = (float)mean;
        stats[2 * b + 1] = (float)sqrt(t / (double)(n > 1 ? n - 1 : 1));
    }
}

__global__ void __launch_bounds__(EW_THREADS)
affine_bcast_kernel(const float* __restrict__ x, float* __restrict__ out, const float* __restrict__ sub, int64_t sub_sb,
                    int64_t sub_si, const float* __restrict__ mul, int64_t mul_sb, int64_t mul_si, int64_t n) {
    const int b = blockIdx.y;
    const float* xa = x + (size_t)b * n;
    float* o = out + (size_t)b * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float v = xa[i];
        if (sub) v -= sub[b * sub_sb + i * sub_si];
        if (mul) v *= mul[b * mul_sb + i * mul_si];
        o[i] = v;
    }
}

__global__ void __launch_bounds__(EW_THREADS)
label_mix_kernel(const float* __restrict__ y, const int64_t* __restrict__ perm, const float* __restrict__ coef,
                 float* __restrict__ out, int64_t n, int hard) {
    const int b = blockIdx.y;
    const float* ya = y + (size_t)b * n;
    const float* yb = y + (size_t)perm[b] * n;
    const float c = coef ? coef[b] : 1.0f;
    float* o = out + (size_t)b * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float v = hard ? ya[i] + yb[i] : c * ya[i] + (1.0f - c) * yb[i];
        o[i] = fminf(fmaxf(v, 0.0f), 1.0f);
    }
}

__global__ void __launch_bounds__(EW_THREADS)
roll_last_kernel(const float* __restrict__ x, float* __restrict__ out, const int32_t* __restrict__ shift, int rows,
                 int cols) {
    const int b = blockIdx.y;
    int s = shift[b] % cols;
    if (s < 0) s += cols;
    const float* xa = x + (size_t)b * rows * cols;
    float* o = out + (size_t)b * rows * cols;
    const int64_t n = (int64_t)rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int r = (int)(i / cols), j = (int)(i - (int64_t)r * cols);
        int src = j - s;
        if (src < 0) src += cols;
        o[i] = xa[(int64_t)r * cols + src];
    }
}

__global__ void __launch_bounds__(EW_THREADS)
add_noise_kernel(const float* __restrict__ x, const float* __restrict__ noise, const float* __restrict__ snr_db,
                 const float* __restrict__ stats, float* __restrict__ out, int64_t n) {
    const int b = blockIdx.y;
    const float sigma = stats[2 * b + 1] / powf(10.0f, snr_db[b] / 20.0f);
    const float* xa = x + (size_t)b * n;
    const float* na = noise + (size_t)b * n;
    float* o = out + (size_t)b * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        o[i] = xa[i] + na[i] * sigma;
}

// fused EMA + Adam on flat buffers (12 B/param EMA traffic + 28 B/param Adam traffic in one pass)
__global__ void __launch_bounds__(EW_THREADS)
adam_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                float* __restrict__ ema, int64_t n, int do_adam, float step_size, float beta1, float beta2, float eps,
                float inv_sqrt_bc2, float ema_alpha, float grad_scale, const float* __restrict__ hyper) {
    pdl_enter();
    if (hyper != nullptr) {
        step_size = hyper[0];
        inv_sqrt_bc2 = hyper[1];
        ema_alpha = hyper[2];
        grad_scale = hyper[3];
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float pi = p[i];
        if (ema) ema[i] = ema_elem(ema[i], pi, ema_alpha);
        if (do_adam) {
            float mi = m[i], vi = v[i];
            adam_elem(pi, g[i], mi, vi, step_size, beta1, beta2, eps, inv_sqrt_bc2, grad_scale);
            m[i] = mi;
            v[i] = vi;
            p[i] = pi;
        }
    }
}

__global__ void __launch_bounds__(EW_THREADS) sumsq_kernel(const float* __restrict__ g, int64_t n, double* out) {
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float v = g[i];
        acc += (double)v * (double)v;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

// median over a window of `k` frames (reflect boundary), one thread per output sample; k <= 31.
// scipy.ndimage.median_filter semantics: window = [i - k/2, i + k - 1 - k/2], rank k/2.
constexpr int MED_MAX = 31;
__global__ void __launch_bounds__(EW_THREADS)
median_kernel(const float* __restrict__ s, float* __restrict__ out, int C, int T, int64_t sb, int64_t sc, int64_t st,
              int64_t ob, int64_t oc, int64_t ot, const int32_t* __restrict__ win, int64_t total) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int t = (int)(i % T);
        int64_t bc = i / T;
        int c = (int)(bc % C);
        int64_t b = bc / C;
        const float* row = s + b * sb + c * sc;
        int k = win[c];
        k = k < 1 ? 1 : (k > MED_MAX ? MED_MAX : k);
        float w[MED_MAX];
        const int half = k / 2;
        for (int j = 0; j < k; j++) {
            int q = t - half + j;
            // scipy 'reflect': (d c b a | a b c d | d c b a)
            int period = 2 * T;
            q %= period;
            if (q < 0) q += period;
            if (q >= T) q = period - 1 - q;
            w[j] = row[(int64_t)q * st];
        }
        // partial selection sort up to rank `half`
        for (int a = 0; a <= half; a++) {
            int mi = a;
            for (int j = a + 1; j < k; j++)
                if (w[j] < w[mi]) mi = j;
            float tmp = w[a];
            w[a] = w[mi];
            w[mi] = tmp;
        }
        out[b * ob + c * oc + (int64_t)t * ot] = w[half];
    }
}

// ---- threshold + run-length event decoding -----------------------------------------------------------------------
// One thread owns one row (threshold th, clip b, class c) and scans its T frames once: an event starts where
// score > threshold turns true and ends at the first frame where it is false again (or at the clip's length).
// COUNT = true: number of events of the row -> cnt[row];  COUNT = false: the row's events -> events[off[row] + e].
template <bool COUNT>
__global__ void __launch_bounds__(EW_THREADS)
decode_events_kernel(const float* __restrict__ s, int B, int C, int T, int64_t sb, int64_t sc, int64_t st,
                     const float* __restrict__ thr, int n_th, const int32_t* __restrict__ n_frames,
                     int32_t* __restrict__ cnt, const int32_t* __restrict__ off, int32_t* __restrict__ events,
                     int capacity) {
    const int rows = n_th * B * C;
    for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < rows; row += gridDim.x * blockDim.x) {
        // rows ordered (th, b, c); neighbouring threads share (th, b) and walk neighbouring classes
        const int c = row % C, b = (row / C) % B, th = row / (C * B);
        const float* p = s + (int64_t)b * sb + (int64_t)c * sc;
        const float t_ = thr[th];
        int len = T;
        if (n_frames != nullptr) len = min(max(n_frames[b], 0), T);
        int n = 0, onset = -1;
        int base = COUNT ? 0 : off[row];
        for (int t = 0; t < len; t++) {
            const bool a = p[(int64_t)t * st] > t_;
            if (a && onset < 0) onset = t;
            if (!a && onset >= 0) {
                if (!COUNT && base + n < capacity) {
                    events[2 * (base + n)] = onset;
                    events[2 * (base + n) + 1] = t;
                }
                n++;
                onset = -1;
            }
        }
        if (onset >= 0) {
            if (!COUNT && base + n < capacity) {
                events[2 * (base + n)] = onset;
                events[2 * (base + n) + 1] = len;
            }
            n++;
        }
        if (COUNT) cnt[row] = n;
    }
}

// exclusive prefix sum of n int32 counts, in place, by ONE block of 1024 threads (n is a few 10^4 rows); v[n] = total
__global__ void __launch_bounds__(1024) exclusive_scan_kernel(int32_t* __restrict__ v, int n) {
    __shared__ int32_t part[1024];
    const int tid = threadIdx.x;
    const int per = (n + 1023) / 1024;
    const int lo = min(tid * per, n), hi = min(lo + per, n);
    int32_t sum = 0;
    for (int i = lo; i < hi; i++) sum += v[i];
    part[tid] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        int32_t add = tid >= o ? part[tid - o] : 0;
        __syncthreads();
        part[tid] += add;
        __syncthreads();
    }
    int32_t run = part[tid] - sum;           // exclusive prefix of this thread's chunk
    for (int i = lo; i < hi; i++) {
        const int32_t c = v[i];
        v[i] = run;
        run += c;
    }
    if (tid == 1023) v[n] = part[1023];
}

// ---- strong-label encoding (ManyHotEncoder.encode_strong_df, desed_task/utils/encoder.py:80-171) ------------------------
// One CTA per clip: labels[b] (C x T, zeroed first) <- for every event of the clip IN ORDER: labels[b][class][onset:offset] =
// value (a later event overwrites an earlier one, as the reference's sequential `y[onset:offset, i] = ...` does).
__global__ void __launch_bounds__(256)
encode_strong_kernel(const int32_t* __restrict__ ev, const float* __restrict__ val, const int32_t* __restrict__ clip_off,
                     float* __restrict__ labels, int C, int T) {
    const int b = blockIdx.x;
    float* y = labels + (size_t)b * C * T;
    for (int i = threadIdx.x; i < C * T; i += blockDim.x) y[i] = 0.f;
    __syncthreads();
    for (int e = clip_off[b]; e < clip_off[b + 1]; e++) {
        const int c = ev[4 * e + 1];
        int on = ev[4 * e + 2], off = ev[4 * e + 3];
        on = on < 0 ? 0 : on;
        off = off > T ? T : off;
        const float v = val != nullptr ? val[e] : 1.0f;
        if (c >= 0 && c < C)
            for (int t = on + (int)threadIdx.x; t < off; t += blockDim.x) y[(size_t)c * T + t] = v;
        __syncthreads();
    }
}

// ---- embedding storage format --------------------------------------------------------------------------
// Time-aggregation of frame embeddings [B, E, Te] -> [B, E, T] with exactly the arithmetic of the fusion kernel
// (emb_concat_kernel: adaptive_avg_pool1d windows summed left to right, or nearest-exact), written as fp32 or bf16.
// Stored this way (768 x 156 bf16 = 240 KB per clip instead of 768 x 496 fp32 = 1.52 MB) the embeddings are read by the
// unchanged fusion path: pooling a [.., 156] input to 156 frames is the identity.
template <bool BF16>
__global__ void __launch_bounds__(256)
pool_embeddings_kernel(const float* __restrict__ emb, void* __restrict__ out, int E, int Te, int T, int mode, int64_t total) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int t = (int)(i % T);
        const int64_t be = i / T;
        const float* ep = emb + be * Te;
        float v;
        if (mode == 1) {
            const float scale = (float)Te / (float)T;
            int q = (int)floorf((float)(((double)t + 0.5) * (double)scale));
            v = ep[q < Te - 1 ? q : Te - 1];
        } else {
            const int st = (int)(((int64_t)t * Te) / T);
            const int en = (int)((((int64_t)(t + 1)) * Te + T - 1) / T);
            float s = 0.f;
            for (int q = st; q < en; q++) s += ep[q];
            v = s / (float)(en - st);
        }
        if (BF16) reinterpret_cast<__nv_bfloat16*>(out)[i] = __float2bfloat16_rn(v);
        else reinterpret_cast<float*>(out)[i] = v;
    }
}

__global__ void __launch_bounds__(256) bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = __bfloat162float(in[i]);
}

inline dim3 grid2(int64_t n, int B, int per_thread = 4) {
    int64_t blocks = (n + (int64_t)EW_THREADS * per_thread - 1) / ((int64_t)EW_THREADS * per_thread);
    int64_t cap = (int64_t)num_sms() * 8 / (B > 0 ? B : 1) + 1;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return dim3((unsigned)blocks, (unsigned)B);
}
inline int grid1(int64_t n, int per_thread = 4) {
    int64_t blocks = (n + (int64_t)EW_THREADS * per_thread - 1) / ((int64_t)EW_THREADS * per_thread);
    int64_t cap = (int64_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace
}  // namespace sedk

using namespace sedk;

extern "C" int sedk_minmax_init(uint32_t* minmax, int B, void* stream) {
    SEDK_REQUIRE(minmax && B > 0, "sedk_minmax_init: bad arguments");
    minmax_init_kernel<<<cdiv(B, 128), 128, 0, (cudaStream_t)stream>>>(minmax, B);
    SEDK_LAUNCH_CHECK("minmax_init_kernel");
    return SEDK_OK;
}

extern "C" int sedk_minmax_decode(const uint32_t* minmax, float* out, int B, void* stream) {
    SEDK_REQUIRE(minmax && out && B > 0, "sedk_minmax_decode: bad arguments");
    minmax_decode_kernel<<<cdiv(2 * B, 128), 128, 0, (cudaStream_t)stream>>>(minmax, out, B);
    SEDK_LAUNCH_CHECK("minmax_decode_kernel");
    return SEDK_OK;
}

extern "C" int sedk_feat_mix_log(const float* x, const int64_t* perm, const float* coef, float* out, int B, int64_t n,
                                 int log_mode, float amin, float db_lo, float db_hi, uint32_t* minmax, void* stream) {
    SEDK_PROF("feat_mix_log", (cudaStream_t)stream);
    SEDK_REQUIRE(x && out && B > 0 && n > 0, "sedk_feat_mix_log: bad arguments");
    const bool vec = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    dim3 grid = grid2(vec ? n / 4 : n, B, 2);
    if (vec)
        feat_mix_log_kernel<true><<<grid, EW_THREADS, 0, (cudaStream_t)stream>>>(x, perm, coef, out, n, log_mode, amin,
                                                                                 db_lo, db_hi, minmax);
    else
        feat_mix_log_kernel<false><<<grid, EW_THREADS, 0, (cudaStream_t)stream>>>(x, perm, coef, out, n, log_mode, amin,
                                                                                  db_lo, db_hi, minmax);
    SEDK_LAUNCH_CHECK("feat_mix_log_kernel");
    return SEDK_OK;
}

extern "C" int sedk_minmax_scale(const float* x, float* out, const uint32_t* minmax, int B, int64_t n, float eps,
                                 void* stream) {
    SEDK_REQUIRE(x && out && minmax && B > 0 && n > 0, "sedk_minmax_scale: bad arguments");
    minmax_scale_kernel<<<grid2(n, B), EW_THREADS, 0, (cudaStream_t)stream>>>(x, out, minmax, n, eps);
    SEDK_LAUNCH_CHECK("minmax_scale_kernel");
    return SEDK_OK;
}

extern "C" int sedk_instance_stats(const float* x, float* stats, int B, int64_t n, void* stream) {
    SEDK_REQUIRE(x && stats && B > 0 && n > 0, "sedk_instance_stats: bad arguments");
    instance_stats_kernel<<<B, EW_THREADS, 0, (cudaStream_t)stream>>>(x, stats, n);
    SEDK_LAUNCH_CHECK("instance_stats_kernel");
    return SEDK_OK;
}

extern "C" int sedk_affine_bcast(const float* x, float* out, const float* sub, int64_t sub_sb, int64_t sub_si,
                                 const float* mul, int64_t mul_sb, int64_t mul_si, int B, int64_t n, void* stream) {
    SEDK_REQUIRE(x && out && B > 0 && n > 0, "sedk_affine_bcast: bad arguments");
    affine_bcast_kernel<<<grid2(n, B), EW_THREADS, 0, (cudaStream_t)stream>>>(x, out, sub, sub_sb, sub_si, mul, mul_sb,
                                                                              mul_si, n);
    SEDK_LAUNCH_CHECK("affine_bcast_kernel");
    return SEDK_OK;
}

extern "C" int sedk_label_mix(const float* y, const int64_t* perm, const float* coef, float* out, int B, int64_t n,
                              int hard, void* stream) {
    SEDK_REQUIRE(y && perm && out && B > 0 && n > 0, "sedk_label_mix: bad arguments");
    label_mix_kernel<<<grid2(n, B), EW_THREADS, 0, (cudaStream_t)stream>>>(y, perm, coef, out, n, hard);
    SEDK_LAUNCH_CHECK("label_mix_kernel");
    return SEDK_OK;
}

extern "C" int sedk_roll_last(const float* x, float* out, const int32_t* shift, int B, int rows, int cols, void* stream) {
    SEDK_REQUIRE(x && out && shift && B > 0 && rows > 0 && cols > 0, "sedk_roll_last: bad arguments");
    SEDK_REQUIRE(x != out, "sedk_roll_last: in-place roll is not supported");
    roll_last_kernel<<<grid2((int64_t)rows * cols, B), EW_THREADS, 0, (cudaStream_t)stream>>>(x, out, shift, rows, cols);
    SEDK_LAUNCH_CHECK("roll_last_kernel");
    return SEDK_OK;
}

extern "C" int sedk_add_noise(const float* x, const float* noise, const float* snr_db, const float* stats, float* out,
                              int B, int64_t n, void* stream) {
    SEDK_REQUIRE(x && noise && snr_db && stats && out && B > 0 && n > 0, "sedk_add_noise: bad arguments");
    add_noise_kernel<<<grid2(n, B), EW_THREADS, 0, (cudaStream_t)stream>>>(x, noise, snr_db, stats, out, n);
    SEDK_LAUNCH_CHECK("add_noise_kernel");
    return SEDK_OK;
}

extern "C" int sedk_adam_ema(float* p, const float* g, float* m, float* v, float* ema, int64_t n, int do_adam, float lr,
                             float beta1, float beta2, float eps, int step, float ema_alpha, float grad_scale,
                             void* stream) {
    SEDK_PROF("adam_ema", (cudaStream_t)stream);
    SEDK_REQUIRE(p && n > 0, "sedk_adam_ema: bad arguments");
    SEDK_REQUIRE(!do_adam || (g && m && v && step >= 1), "sedk_adam_ema: Adam needs g, m, v and step >= 1");
    float step_size = 0.f, inv_sqrt_bc2 = 1.f;
    if (do_adam) {
        double bc1 = 1.0 - pow((double)beta1, (double)step);
        double bc2 = 1.0 - pow((double)beta2, (double)step);
        step_size = (float)((double)lr / bc1);
        inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    }
    SEDK_CUDA(pdl_launch(adam_ema_kernel, dim3(grid1(n, 1)), dim3(EW_THREADS), (size_t)(0), (cudaStream_t)stream, p, g, m, v, ema, n, do_adam, step_size, beta1,
                                                                        beta2, eps, inv_sqrt_bc2, ema_alpha, grad_scale, nullptr));
    SEDK_LAUNCH_CHECK("adam_ema_kernel");
    return SEDK_OK;
}

__global__ void bump_kernel(uint64_t* c, uint64_t inc) { *c += inc; }

extern "C" int sedk_adam_ema_dev(float* p, const float* g, float* m, float* v, float* ema, int64_t n, int do_adam,
                                 float beta1, float beta2, float eps, const float* hyper, void* stream) {
    SEDK_PROF("adam_ema", (cudaStream_t)stream);
    SEDK_REQUIRE(p && hyper && n > 0, "sedk_adam_ema_dev: bad arguments");
    SEDK_REQUIRE(!do_adam || (g && m && v), "sedk_adam_ema_dev: Adam needs g, m, v");
    SEDK_CUDA(pdl_launch(adam_ema_kernel, dim3(grid1(n, 1)), dim3(EW_THREADS), (size_t)(0), (cudaStream_t)stream, p, g, m, v, ema, n, do_adam, 0.f, beta1, beta2,
                                                                        eps, 1.f, 0.f, 1.f, hyper));
    SEDK_LAUNCH_CHECK("adam_ema_kernel");
    return SEDK_OK;
}

extern "C" int sedk_bump_counter(uint64_t* counter, uint64_t inc, void* stream) {
    SEDK_REQUIRE(counter, "sedk_bump_counter: null counter");
    bump_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter, inc);
    SEDK_LAUNCH_CHECK("bump_kernel");
    return SEDK_OK;
}

// torchaudio mask_along_axis_iid draws (functional.py:857-869) for two masked axes at once, on the device:
//   value = u * param;  min = u' * (size - value);  span = [floor(min), floor(min) + floor(value))
// out[b] = {start_a, end_a, start_b, end_b}; param < 1 disables an axis (span 0, 0).  One Philox call per example.
namespace sedk {
namespace {
__global__ void mask_spans_kernel(int32_t* __restrict__ out, int B, int size_a, int param_a, int size_b, int param_b,
                                  uint64_t seed, const uint64_t* __restrict__ seed_dev, uint64_t stream_id) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const Philox ph(seed + (seed_dev ? *seed_dev : 0ull));
    const uint4 r = ph((uint64_t)b, stream_id);
    const float k = 2.3283064365386963e-10f;     // 2^-32
    int sa = 0, ea = 0, sb = 0, eb = 0;
    if (param_a >= 1) {
        const float value = (float)r.x * k * (float)param_a;
        const float mn = (float)r.y * k * ((float)size_a - value);
        sa = (int)mn;
        ea = sa + (int)value;
    }
    if (param_b >= 1) {
        const float value = (float)r.z * k * (float)param_b;
        const float mn = (float)r.w * k * ((float)size_b - value);
        sb = (int)mn;
        eb = sb + (int)value;
    }
    reinterpret_cast<int4*>(out)[b] = make_int4(sa, ea, sb, eb);
}
}  // namespace
}  // namespace sedk

extern "C" int sedk_mask_spans(int32_t* out, int B, int size_a, int param_a, int size_b, int param_b, uint64_t seed,
                               const uint64_t* seed_dev, uint64_t stream_id, void* stream) {
    SEDK_REQUIRE(out && B > 0, "sedk_mask_spans: bad arguments");
    sedk::mask_spans_kernel<<<sedk::cdiv(B, 128), 128, 0, (cudaStream_t)stream>>>(out, B, size_a, param_a, size_b, param_b,
                                                                                  seed, seed_dev, stream_id);
    SEDK_LAUNCH_CHECK("mask_spans_kernel");
    return SEDK_OK;
}

extern "C" int sedk_sumsq(const float* g, int64_t n, double* out, void* stream) {
    SEDK_REQUIRE(g && out && n > 0, "sedk_sumsq: bad arguments");
    SEDK_CUDA(cudaMemsetAsync(out, 0, sizeof(double), (cudaStream_t)stream));
    sumsq_kernel<<<grid1(n), EW_THREADS, 0, (cudaStream_t)stream>>>(g, n, out);
    SEDK_LAUNCH_CHECK("sumsq_kernel");
    return SEDK_OK;
}

extern "C" int sedk_median_filter(const float* scores, float* out, int B, int C, int T, int64_t sb, int64_t sc,
                                  int64_t st, int64_t ob, int64_t oc, int64_t ot, const int32_t* win, void* stream) {
    SEDK_REQUIRE(scores && out && win && B > 0 && C > 0 && T > 0, "sedk_median_filter: bad arguments");
    SEDK_REQUIRE(scores != out, "sedk_median_filter: in-place filtering is not supported");
    int64_t total = (int64_t)B * C * T;
    median_kernel<<<grid1(total, 1), EW_THREADS, 0, (cudaStream_t)stream>>>(scores, out, C, T, sb, sc, st, ob, oc, ot, win,
                                                                           total);
    SEDK_LAUNCH_CHECK("median_kernel");
    return SEDK_OK;
}

extern "C" int sedk_decode_events(const float* scores, int B, int C, int T, int64_t sb, int64_t sc, int64_t st,
                                  const float* thresholds, int n_th, const int32_t* n_frames, int32_t* offsets,
                                  int32_t* events, int capacity, void* stream) {
    using namespace sedk;
    SEDK_REQUIRE(scores && thresholds && offsets && events, "sedk_decode_events: null pointer");
    SEDK_REQUIRE(B > 0 && C > 0 && T > 0 && n_th > 0 && capacity >= 0, "sedk_decode_events: bad sizes");
    SEDK_REQUIRE((int64_t)n_th * B * C < (1ll << 30), "sedk_decode_events: too many rows");
    cudaStream_t s = (cudaStream_t)stream;
    SEDK_PROF("decode_events", s);
    const int rows = n_th * B * C;
    const int blocks = cdiv(rows, EW_THREADS);
    decode_events_kernel<true><<<blocks, EW_THREADS, 0, s>>>(scores, B, C, T, sb, sc, st, thresholds, n_th, n_frames,
                                                            offsets, nullptr, nullptr, 0);
    SEDK_LAUNCH_CHECK("decode_events_kernel<count>");
    exclusive_scan_kernel<<<1, 1024, 0, s>>>(offsets, rows);
    SEDK_LAUNCH_CHECK("exclusive_scan_kernel");
    decode_events_kernel<false><<<blocks, EW_THREADS, 0, s>>>(scores, B, C, T, sb, sc, st, thresholds, n_th, n_frames,
                                                             nullptr, offsets, events, capacity);
    SEDK_LAUNCH_CHECK("decode_events_kernel<write>");
    return SEDK_OK;
}

extern "C" int sedk_encode_strong(const int32_t* events, const float* values, const int32_t* clip_offsets, float* labels, int B,
                                  int C, int T, void* stream) {
    using namespace sedk;
    SEDK_REQUIRE(events && clip_offsets && labels && B > 0 && C > 0 && T > 0, "sedk_encode_strong: bad arguments");
    SEDK_PROF("encode_strong", (cudaStream_t)stream);
    encode_strong_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(events, values, clip_offsets, labels, C, T);
    SEDK_LAUNCH_CHECK("encode_strong_kernel");
    return SEDK_OK;
}

extern "C" int sedk_pool_embeddings(const float* emb, void* out, int B, int E, int Te, int T, int mode, int out_bf16,
                                    void* stream) {
    using namespace sedk;
    SEDK_REQUIRE(emb && out && B > 0 && E > 0 && Te > 0 && T > 0 && (mode == 0 || mode == 1), "sedk_pool_embeddings: bad arguments");
    SEDK_PROF("pool_embeddings", (cudaStream_t)stream);
    const int64_t total = (int64_t)B * E * T;
    if (out_bf16)
        pool_embeddings_kernel<true><<<grid1(total, 1), EW_THREADS, 0, (cudaStream_t)stream>>>(emb, out, E, Te, T, mode, total);
    else
        pool_embeddings_kernel<false><<<grid1(total, 1), EW_THREADS, 0, (cudaStream_t)stream>>>(emb, out, E, Te, T, mode, total);
    SEDK_LAUNCH_CHECK("pool_embeddings_kernel");
    return SEDK_OK;
}

extern "C" int sedk_bf16_to_f32(const void* in, float* out, int64_t n, void* stream) {
    using namespace sedk;
    SEDK_REQUIRE(in && out && n > 0, "sedk_bf16_to_f32: bad arguments");
    SEDK_PROF("bf16_to_f32", (cudaStream_t)stream);
    bf16_to_f32_kernel<<<grid1(n, 4), EW_THREADS, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(in), out, n);
    SEDK_LAUNCH_CHECK("bf16_to_f32_kernel");
    return SEDK_OK;
}
