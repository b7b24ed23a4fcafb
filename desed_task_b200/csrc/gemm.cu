// Plain tensor-core GEMM building block (TF32 / 3xTF32 mma.sync) used for the GRU input projections, their
// gradients, and the embedding-fusion linear layer.  C[M,N] = alpha * op(A) op(B) + beta * C + bias[n].
// 64x64x32 CTA tile, 4 warps (32x32 each); arbitrary M, N, K and leading dimensions (bounds-checked scalar loads);
// split-K with atomic accumulation for the tall-skinny weight-gradient shapes (K = B*T).
#include "kernels.h"

namespace sedk {
namespace {

constexpr int BM = 64, BN = 64, BK = 32;

template <bool TA, bool TB, bool X3>
__global__ void __launch_bounds__(128)
gemm_kernel(int M, int N, int K, float alpha, const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb,
            float beta, float* __restrict__ C, int ldc, const float* __restrict__ bias, int k_per_split, int atomic) {
    // A tile: !TA -> As[m][k] (stride BK+4), TA -> As[k][m] (stride BM+8);  B tile: TB -> Bs[n][k], !TB -> Bs[k][n]
    constexpr int ASZ = TA ? BK * (BM + 8) : BM * (BK + 4);
    constexpr int BSZ = TB ? BN * (BK + 4) : BK * (BN + 8);
    __shared__ float As[ASZ];
    __shared__ float Bs[BSZ];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int wm0 = (warp >> 1) * 32, wn0 = (warp & 1) * 32;
    const int kbeg = blockIdx.z * k_per_split;
    const int kend = min(K, kbeg + k_per_split);
    float acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int q = 0; q < 4; q++) acc[i][j][q] = 0.f;

    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        // ---- stage A
        if (!TA) {
            for (int idx = tid; idx < BM * BK; idx += 128) {
                int m = idx / BK, k = idx - m * BK;
                int gm = m0 + m, gk = k0 + k;
                As[m * (BK + 4) + k] = (gm < M && gk < kend) ? A[(size_t)gm * lda + gk] : 0.f;
            }
        } else {
            for (int idx = tid; idx < BM * BK; idx += 128) {
                int k = idx / BM, m = idx - k * BM;
                int gm = m0 + m, gk = k0 + k;
                As[k * (BM + 8) + m] = (gm < M && gk < kend) ? A[(size_t)gk * lda + gm] : 0.f;
            }
        }
        // ---- stage B
        if (TB) {
            for (int idx = tid; idx < BN * BK; idx += 128) {
                int n = idx / BK, k = idx - n * BK;
                int gn = n0 + n, gk = k0 + k;
                Bs[n * (BK + 4) + k] = (gn < N && gk < kend) ? Bm[(size_t)gn * ldb + gk] : 0.f;
            }
        } else {
            for (int idx = tid; idx < BN * BK; idx += 128) {
                int k = idx / BN, n = idx - k * BN;
                int gn = n0 + n, gk = k0 + k;
                Bs[k * (BN + 8) + n] = (gn < N && gk < kend) ? Bm[(size_t)gk * ldb + gn] : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k8 = 0; k8 < BK / 8; k8++) {
            auto fa = [&](int i, int r, int c) {
                const int m = wm0 + i * 16 + g + 8 * r, k = k8 * 8 + t4 + 4 * c;
                return TA ? As[k * (BM + 8) + m] : As[m * (BK + 4) + k];
            };
            auto fb = [&](int j, int c) {
                const int n = wn0 + j * 8 + g, k = k8 * 8 + t4 + 4 * c;
                return TB ? Bs[n * (BK + 4) + k] : Bs[k * (BN + 8) + n];
            };
            warp_mma_k8<2, 4, X3>(acc, fa, fb);
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
            for (int j = 0; j < 4; j++)
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    const int n = n0 + wn0 + j * 8 + 2 * t4 + q;
                    if (n >= N) continue;
                    float v = alpha * acc[i][j][2 * r + q];
                    float* cp = C + (size_t)m * ldc + n;
                    if (atomic) {
                        if (blockIdx.z == 0 && bias != nullptr) v += bias[n];
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

}  // namespace

int launch_gemm(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda, const float* B,
                int ldb, float beta, float* C, int ldc, const float* bias, int precision, cudaStream_t s) {
    char pname[64];
    snprintf(pname, sizeof(pname), "gemm_%c%c_%dx%dx%d", transA ? 'T' : 'N', transB ? 'T' : 'N', M, N, K);
    SEDK_PROF(pname, s);
    SEDK_REQUIRE(M > 0 && N > 0 && K > 0 && A && B && C, "gemm: bad arguments");
    dim3 grid(cdiv(N, BN), cdiv(M, BM), 1);
    int k_per_split = cdiv(K, BK) * BK;
    int atomic = 0;
    const int tiles = grid.x * grid.y;
    if (beta == 1.0f && K >= 1024 && tiles < num_sms()) {
        // accumulate-into-C shapes (weight gradients): split K and add atomically
        int splits = min(cdiv(2 * num_sms(), tiles), cdiv(K, 4 * BK));
        if (splits > 1) {
            k_per_split = cdiv(cdiv(K, splits), BK) * BK;
            grid.z = cdiv(K, k_per_split);
            atomic = 1;
        }
    }
#define SEDK_GEMM(TA, TB)                                                                                              \
    if (precision)                                                                                                     \
        gemm_kernel<TA, TB, true><<<grid, 128, 0, s>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, k_per_split, \
                                                       atomic);                                                       \
    else                                                                                                               \
        gemm_kernel<TA, TB, false><<<grid, 128, 0, s>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, k_per_split, \
                                                        atomic);
    if (transA && transB) { SEDK_GEMM(true, true) }
    else if (transA) { SEDK_GEMM(true, false) }
    else if (transB) { SEDK_GEMM(false, true) }
    else { SEDK_GEMM(false, false) }
#undef SEDK_GEMM
    SEDK_LAUNCH_CHECK("gemm_kernel");
    return SEDK_OK;
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
