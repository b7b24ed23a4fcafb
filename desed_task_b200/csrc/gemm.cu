// Plain tensor-core GEMM building block (TF32 / 3xTF32 mma.sync) for the GRU input projections, their gradients and
// the embedding-fusion linear layer:  C[M,N] = alpha * op(A) op(B) + beta * C + bias[n], up to 4 same-shape problems
// per launch (blockIdx.z) so that the two GRU directions share one launch.
//
// Fast path: 128x64x32 CTA tile, 8 warps (32x32 each), operands staged by 16-byte cp.async into a double-buffered
// shared-memory ring.  Both storage orders of each operand keep their contiguous axis contiguous in shared memory
// (A: [m][k] or [k][m]; B: [n][k] or [k][n]) with paddings chosen so that every mma fragment load is bank-conflict free.
// Split-K with atomic accumulation serves the tall-skinny weight-gradient shapes (K = B*T).  Shapes or pointers that
// break the 16-byte alignment rules take the bounds-checked scalar staging path of the same kernel.
#include "kernels.h"

namespace sedk {
namespace {

constexpr int BK = 32;

struct GemmBatch {
    const float* A[4];
    const float* B[4];
    float* C[4];
    const float* bias[4];
};

template <int BM, int BN, bool TA, bool TB>
struct Tile {
    static constexpr int AS = TA ? BM + 8 : BK + 4;           // row stride of the A tile
    static constexpr int BS = TB ? BK + 4 : BN + 8;
    static constexpr int ASZ = TA ? BK * AS : BM * AS;
    static constexpr int BSZ = TB ? BN * BS : BK * BS;
};

template <int BM, int BN, bool TA, bool TB, bool X3, bool VEC>
__global__ void __launch_bounds__(256)
gemm_kernel(int M, int N, int K, float alpha, GemmBatch gb, int lda, int ldb, float beta, int ldc, int k_per_split,
            int n_split, int atomic) {
    using Tl = Tile<BM, BN, TA, TB>;
    constexpr int AS = Tl::AS, BS = Tl::BS, ASZ = Tl::ASZ, BSZ = Tl::BSZ;
    constexpr int WM = BM / 32, WN = 8 / WM;                   // 8 warps
    constexpr int NF = BN / WN / 8;
    extern __shared__ float smem[];
    float* As = smem;                 // [2][ASZ]
    float* Bs = smem + 2 * ASZ;       // [2][BSZ]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int prob = blockIdx.z / n_split, split = blockIdx.z - prob * n_split;
    const float* __restrict__ A = gb.A[prob];
    const float* __restrict__ Bm = gb.B[prob];
    float* __restrict__ C = gb.C[prob];
    const float* __restrict__ bias = gb.bias[prob];
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int wm0 = (warp / WN) * 32, wn0 = (warp % WN) * (NF * 8);
    const int kbeg = split * k_per_split;
    const int kend = min(K, kbeg + k_per_split);
    float acc[2][NF][4];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < NF; j++)
#pragma unroll
            for (int q = 0; q < 4; q++) acc[i][j][q] = 0.f;

    auto stage = [&](int k0, int buf) {
        float* as = As + buf * ASZ;
        float* bs = Bs + buf * BSZ;
        if (VEC) {
            if (!TA) {
                for (int idx = tid; idx < BM * (BK / 4); idx += 256) {
                    int m = idx / (BK / 4), q = idx - m * (BK / 4);
                    int gm = m0 + m, gk = k0 + q * 4;
                    bool ok = gm < M && gk < kend;
                    cp_async16(as + m * AS + q * 4, ok ? A + (size_t)gm * lda + gk : A, ok);
                }
            } else {
                for (int idx = tid; idx < BK * (BM / 4); idx += 256) {
                    int k = idx / (BM / 4), q = idx - k * (BM / 4);
                    int gm = m0 + q * 4, gk = k0 + k;
                    bool ok = gm < M && gk < kend;
                    cp_async16(as + k * AS + q * 4, ok ? A + (size_t)gk * lda + gm : A, ok);
                }
            }
            if (TB) {
                for (int idx = tid; idx < BN * (BK / 4); idx += 256) {
                    int n = idx / (BK / 4), q = idx - n * (BK / 4);
                    int gn = n0 + n, gk = k0 + q * 4;
                    bool ok = gn < N && gk < kend;
                    cp_async16(bs + n * BS + q * 4, ok ? Bm + (size_t)gn * ldb + gk : Bm, ok);
                }
            } else {
                for (int idx = tid; idx < BK * (BN / 4); idx += 256) {
                    int k = idx / (BN / 4), q = idx - k * (BN / 4);
                    int gn = n0 + q * 4, gk = k0 + k;
                    bool ok = gn < N && gk < kend;
                    cp_async16(bs + k * BS + q * 4, ok ? Bm + (size_t)gk * ldb + gn : Bm, ok);
                }
            }
        } else {
            for (int idx = tid; idx < BM * BK; idx += 256) {
                int m, k;
                if (!TA) { m = idx / BK; k = idx - m * BK; } else { k = idx / BM; m = idx - k * BM; }
                int gm = m0 + m, gk = k0 + k;
                float v = (gm < M && gk < kend) ? (TA ? A[(size_t)gk * lda + gm] : A[(size_t)gm * lda + gk]) : 0.f;
                as[TA ? k * AS + m : m * AS + k] = v;
            }
            for (int idx = tid; idx < BN * BK; idx += 256) {
                int n, k;
                if (TB) { n = idx / BK; k = idx - n * BK; } else { k = idx / BN; n = idx - k * BN; }
                int gn = n0 + n, gk = k0 + k;
                float v = (gn < N && gk < kend) ? (TB ? Bm[(size_t)gn * ldb + gk] : Bm[(size_t)gk * ldb + gn]) : 0.f;
                bs[TB ? n * BS + k : k * BS + n] = v;
            }
        }
    };

    const int nk = (kend - kbeg + BK - 1) / BK;
    if (nk > 0) stage(kbeg, 0);
    cp_async_commit();
    for (int it = 0; it < nk; it++) {
        if (it + 1 < nk) stage(kbeg + (it + 1) * BK, (it + 1) & 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const float* as = As + (it & 1) * ASZ;
        const float* bs = Bs + (it & 1) * BSZ;
#pragma unroll
        for (int k8 = 0; k8 < BK / 8; k8++) {
            auto fa = [&](int i, int r, int c) {
                const int m = wm0 + i * 16 + g + 8 * r, k = k8 * 8 + t4 + 4 * c;
                return TA ? as[k * AS + m] : as[m * AS + k];
            };
            auto fb = [&](int j, int c) {
                const int n = wn0 + j * 8 + g, k = k8 * 8 + t4 + 4 * c;
                return TB ? bs[n * BS + k] : bs[k * BS + n];
            };
            warp_mma_k8<2, NF, X3>(acc, fa, fb);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const int m = m0 + wm0 + i * 16 + g + 8 * r;
            if (m >= M) continue;
#pragma unroll
            for (int j = 0; j < NF; j++)
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    const int n = n0 + wn0 + j * 8 + 2 * t4 + q;
                    if (n >= N) continue;
                    float v = alpha * acc[i][j][2 * r + q];
                    float* cp = C + (size_t)m * ldc + n;
                    if (atomic) {
                        if (split == 0 && bias != nullptr) v += bias[n];
                        atomicAdd(cp, v);
                    } else {
                        if (bias != nullptr) v += bias[n];
                        if (beta != 0.f) v += beta * (*cp);
                        *cp = v;
                    }
                }
        }
}

__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ A, int M, int N, int lda, float* __restrict__ out, int rows_per_block) {
    // block = 32 columns x 8 row-lanes
    __shared__ float red[8][33];
    const int col = blockIdx.x * 32 + (threadIdx.x & 31);
    const int rl = threadIdx.x >> 5;
    const int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
    float acc = 0.f;
    if (col < N)
        for (int r = r0 + rl; r < r1; r += 8) acc += A[(size_t)r * lda + col];
    red[rl][threadIdx.x & 31] = acc;
    __syncthreads();
    if (rl == 0 && col < N) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) s += red[i][threadIdx.x & 31];
        atomicAdd(&out[col], s);
    }
}

template <int BM, int BN, bool TA, bool TB>
int run_gemm(int M, int N, int K, float alpha, const GemmBatch& gb, int nprob, int lda, int ldb, float beta, int ldc,
             int precision, bool vec, cudaStream_t s) {
    using Tl = Tile<BM, BN, TA, TB>;
    const size_t smem = (size_t)2 * (Tl::ASZ + Tl::BSZ) * sizeof(float);
    dim3 grid(cdiv(N, BN), cdiv(M, BM), 1);
    int k_per_split = cdiv(K, BK) * BK;
    int n_split = 1, atomic = 0;
    const int tiles = grid.x * grid.y * nprob;
    if (beta == 1.0f && K >= 1024 && tiles < num_sms()) {
        int splits = min(cdiv(2 * num_sms(), tiles), cdiv(K, 4 * BK));
        if (splits > 1) {
            k_per_split = cdiv(cdiv(K, splits), BK) * BK;
            n_split = cdiv(K, k_per_split);
            atomic = 1;
        }
    }
    grid.z = nprob * n_split;
#define SEDK_GEMM_LAUNCH(X3, VEC)                                                                                       \
    {                                                                                                                   \
        auto kern = gemm_kernel<BM, BN, TA, TB, X3, VEC>;                                                               \
        static bool cfg = false;                                                                                        \
        if (!cfg) { int rc = opt_in_smem(kern, smem); if (rc) return rc; cfg = true; }                                  \
        kern<<<grid, 256, smem, s>>>(M, N, K, alpha, gb, lda, ldb, beta, ldc, k_per_split, n_split, atomic);            \
    }
    if (precision) { if (vec) SEDK_GEMM_LAUNCH(true, true) else SEDK_GEMM_LAUNCH(true, false) }
    else { if (vec) SEDK_GEMM_LAUNCH(false, true) else SEDK_GEMM_LAUNCH(false, false) }
#undef SEDK_GEMM_LAUNCH
    SEDK_LAUNCH_CHECK("gemm_kernel");
    return SEDK_OK;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

int launch_gemm_batched(int transA, int transB, int M, int N, int K, float alpha, const float* const* A, int lda,
                        const float* const* B, int ldb, float beta, float* const* C, int ldc, const float* const* bias,
                        int nprob, int precision, cudaStream_t s) {
    char pname[64];
    snprintf(pname, sizeof(pname), "gemm_%c%c_%dx%dx%d_x%d", transA ? 'T' : 'N', transB ? 'T' : 'N', M, N, K, nprob);
    SEDK_PROF(pname, s);
    SEDK_REQUIRE(M > 0 && N > 0 && K > 0 && nprob >= 1 && nprob <= 4, "gemm: bad arguments");
    GemmBatch gb;
    bool vec = (lda % 4 == 0) && (ldb % 4 == 0) && (K % 4 == 0);
    if (transA) vec = vec && (M % 4 == 0);
    if (!transB) vec = vec && (N % 4 == 0);
    for (int i = 0; i < 4; i++) {
        int j = i < nprob ? i : 0;
        SEDK_REQUIRE(A[j] && B[j] && C[j], "gemm: null operand");
        gb.A[i] = A[j];
        gb.B[i] = B[j];
        gb.C[i] = C[j];
        gb.bias[i] = bias ? bias[j] : nullptr;
        vec = vec && aligned16(A[j]) && aligned16(B[j]);
    }
#define SEDK_GEMM_DISPATCH(TA, TB) \
    return run_gemm<128, 64, TA, TB>(M, N, K, alpha, gb, nprob, lda, ldb, beta, ldc, precision, vec, s);
    if (transA && transB) { SEDK_GEMM_DISPATCH(true, true) }
    else if (transA) { SEDK_GEMM_DISPATCH(true, false) }
    else if (transB) { SEDK_GEMM_DISPATCH(false, true) }
    else { SEDK_GEMM_DISPATCH(false, false) }
#undef SEDK_GEMM_DISPATCH
}

int launch_gemm(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda, const float* B,
                int ldb, float beta, float* C, int ldc, const float* bias, int precision, cudaStream_t s) {
    SEDK_REQUIRE(A && B && C, "gemm: null operand");
    const float* As[1] = {A};
    const float* Bs[1] = {B};
    float* Cs[1] = {C};
    const float* bs[1] = {bias};
    return launch_gemm_batched(transA, transB, M, N, K, alpha, As, lda, Bs, ldb, beta, Cs, ldc, bias ? bs : nullptr, 1,
                               precision, s);
}

int launch_colsum(const float* A, int M, int N, int lda, float* out, int accumulate, cudaStream_t s) {
    SEDK_PROF("colsum", s);
    if (!accumulate) SEDK_CUDA(cudaMemsetAsync(out, 0, (size_t)N * sizeof(float), s));
    int splits = min(cdiv(M, 64), max(1, 2 * num_sms() / cdiv(N, 32)));
    int rpb = cdiv(M, splits);
    dim3 grid(cdiv(N, 32), cdiv(M, rpb));
    colsum_kernel<<<grid, 256, 0, s>>>(A, M, N, lda, out, rpb);
    SEDK_LAUNCH_CHECK("colsum_kernel");
    return SEDK_OK;
}

}  // namespace sedk

extern "C" int sedk_gemm(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda,
                         const float* Bm, int ldb, float beta, float* C, int ldc, const float* bias, int precision,
                         void* stream) {
    return sedk::launch_gemm(transA, transB, M, N, K, alpha, A, lda, Bm, ldb, beta, C, ldc, bias, precision,
                             (cudaStream_t)stream);
}
