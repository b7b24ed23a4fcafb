// Micro-benchmarks of the SM-level constants the GRU recurrence design depends on (B200, sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu ; prints cycles per warp-instruction.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// mode 0: HMMA tf32, CH independent accumulator chains per warp; 1: HMMA bf16
template <int CH, int MODE>
__global__ void k_mma(long long* out, float* sink, int iters) {
    float d[CH][4];
    uint32_t a[4] = {0x3f800000u + threadIdx.x, 0x3f000000u, 0x3e800000u, 0x3f800000u}, b[2] = {0x3f800000u, 0x3f000000u};
#pragma unroll
    for (int c = 0; c < CH; c++) d[c][0] = d[c][1] = d[c][2] = d[c][3] = 0.f;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CH; c++) {
            if (MODE == 0) mma_tf32(d[c], a, b); else mma_bf16(d[c], a, b);
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CH; c++) s += d[c][0] + d[c][1] + d[c][2] + d[c][3];
    if (s == 12345.f) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int CH>
__global__ void k_ffma2(long long* out, float* sink, int iters) {
    float2 acc[CH];
    float2 w = make_float2(1.0001f, 0.9999f), h = make_float2(0.5f + threadIdx.x * 1e-6f, 0.25f);
#pragma unroll
    for (int c = 0; c < CH; c++) acc[c] = make_float2(0.f, 0.f);
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CH; c++) acc[c] = __ffma2_rn(w, h, acc[c]);
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CH; c++) s += acc[c].x + acc[c].y;
    if (s == 12345.f) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int CH>
__global__ void k_ffma(long long* out, float* sink, int iters) {
    float acc[CH];
    float w = 1.0001f, h = 0.5f + threadIdx.x * 1e-6f;
#pragma unroll
    for (int c = 0; c < CH; c++) acc[c] = 0.f;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CH; c++) acc[c] = fmaf(w, h, acc[c]);
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CH; c++) s += acc[c];
    if (s == 12345.f) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

// pattern 0: all lanes same 16 B; 1: quad pattern (lane&3 selects one of 4 chunks, padded 36 floats apart);
// 2: octet pattern (lane&7 -> 8 chunks, 20 floats apart); 3: every lane its own 16 B (conflict-free); 4: LDS.32 all same;
// 5: LDS.64 all lanes same
template <int PAT>
__global__ void k_lds(long long* out, float* sink, int iters) {
    __shared__ __align__(16) float sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int off = 0;
    if (PAT == 1) off = 36 * (lane & 3);
    if (PAT == 2) off = 20 * (lane & 7);
    if (PAT == 3) off = 4 * lane;
    float4 acc = make_float4(0, 0, 0, 0);
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < 8; c++) {
            if (PAT == 4) {
                float v = *reinterpret_cast<volatile float*>(sm + ((c * 4 + i) & 1023));
                acc.x += v;
            } else if (PAT == 5) {
                float2 v;
                const uint32_t addr = (uint32_t)__cvta_generic_to_shared(sm + ((c * 4 + 2 * i) & 1022));
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];\n" : "=f"(v.x), "=f"(v.y) : "r"(addr));
                acc.x += v.x; acc.y += v.y;
            } else {
                float4 v;
                const uint32_t addr = (uint32_t)__cvta_generic_to_shared(sm + off + ((4 * c + 32 * i) & 1023));
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
    }
    long long t1 = clock64();
    if (acc.x + acc.y + acc.z + acc.w == 12345.f) sink[0] = acc.x;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int CH>
__global__ void k_shfl(long long* out, float* sink, int iters) {
    float v[CH];
#pragma unroll
    for (int c = 0; c < CH; c++) v[c] = threadIdx.x + c;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CH; c++) v[c] += __shfl_xor_sync(0xffffffffu, v[c], 1 + (i & 3));
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int c = 0; c < CH; c++) s += v[c];
    if (s == 12345.f) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int CH>
__global__ void k_mufu(long long* out, float* sink, int iters) {
    float v[CH];
#pragma unroll
    for (int c = 0; c < CH; c++) v[c] = 0.001f * (threadIdx.x + c);
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CH; c++) {
            float y;
            asm volatile("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(v[c]));
            asm volatile("rcp.approx.ftz.f32 %0, %1;\n" : "=f"(v[c]) : "f"(y));
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int c = 0; c < CH; c++) s += v[c];
    if (s == 12345.f) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

__global__ void k_bar(long long* out, int iters) {
    __shared__ float x[1024];
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        x[threadIdx.x] = i;
        __syncthreads();
        if (x[(threadIdx.x + 33) & (blockDim.x - 1)] < 0) out[1] = 1;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <class F>
double run(F launch, int per_iter, int iters) {
    long long* d;
    cudaMalloc(&d, 16);
    launch(d, 8);        // warm
    launch(d, iters);
    cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    return (double)h / ((double)iters * per_iter);
}

int main() {
    float* sink;
    cudaMalloc(&sink, 16);
    const int IT = 2000;
    printf("cycles per warp-instruction as seen by ONE warp (1 CTA on 1 SM); T = threads per CTA\n");
    for (int T : {32, 128, 256, 512}) {
        printf("T=%4d | HMMA.tf32 m16n8k8: ch1 %.2f ch2 %.2f ch4 %.2f ch8 %.2f | HMMA.bf16 m16n8k16: ch1 %.2f ch4 %.2f ch8 %.2f\n", T,
               run([&](long long* d, int it) { k_mma<1, 0><<<1, T>>>(d, sink, it); }, 1, IT),
               run([&](long long* d, int it) { k_mma<2, 0><<<1, T>>>(d, sink, it); }, 2, IT),
               run([&](long long* d, int it) { k_mma<4, 0><<<1, T>>>(d, sink, it); }, 4, IT),
               run([&](long long* d, int it) { k_mma<8, 0><<<1, T>>>(d, sink, it); }, 8, IT),
               run([&](long long* d, int it) { k_mma<1, 1><<<1, T>>>(d, sink, it); }, 1, IT),
               run([&](long long* d, int it) { k_mma<4, 1><<<1, T>>>(d, sink, it); }, 4, IT),
               run([&](long long* d, int it) { k_mma<8, 1><<<1, T>>>(d, sink, it); }, 8, IT));
        printf("T=%4d | FFMA2: ch1 %.2f ch2 %.2f ch4 %.2f ch8 %.2f | FFMA: ch1 %.2f ch4 %.2f ch8 %.2f\n", T,
               run([&](long long* d, int it) { k_ffma2<1><<<1, T>>>(d, sink, it); }, 1, IT),
               run([&](long long* d, int it) { k_ffma2<2><<<1, T>>>(d, sink, it); }, 2, IT),
               run([&](long long* d, int it) { k_ffma2<4><<<1, T>>>(d, sink, it); }, 4, IT),
               run([&](long long* d, int it) { k_ffma2<8><<<1, T>>>(d, sink, it); }, 8, IT),
               run([&](long long* d, int it) { k_ffma<1><<<1, T>>>(d, sink, it); }, 1, IT),
               run([&](long long* d, int it) { k_ffma<4><<<1, T>>>(d, sink, it); }, 4, IT),
               run([&](long long* d, int it) { k_ffma<8><<<1, T>>>(d, sink, it); }, 8, IT));
        printf("T=%4d | LDS.128: same-addr %.2f quad %.2f octet %.2f distinct %.2f | LDS.32 same %.2f | LDS.64 same %.2f\n", T,
               run([&](long long* d, int it) { k_lds<0><<<1, T>>>(d, sink, it); }, 8, IT),
               run([&](long long* d, int it) { k_lds<1><<<1, T>>>(d, sink, it); }, 8, IT),
               run([&](long long* d, int it) { k_lds<2><<<1, T>>>(d, sink, it); }, 8, IT),
               run([&](long long* d, int it) { k_lds<3><<<1, T>>>(d, sink, it); }, 8, IT),
               run([&](long long* d, int it) { k_lds<4><<<1, T>>>(d, sink, it); }, 8, IT),
               run([&](long long* d, int it) { k_lds<5><<<1, T>>>(d, sink, it); }, 8, IT));
        printf("T=%4d | SHFL+FADD: ch1 %.2f ch4 %.2f ch8 %.2f | MUFU ex2+rcp pair: ch1 %.2f ch4 %.2f ch8 %.2f | STS+BAR+LDS %.2f\n", T,
               run([&](long long* d, int it) { k_shfl<1><<<1, T>>>(d, sink, it); }, 1, IT),
               run([&](long long* d, int it) { k_shfl<4><<<1, T>>>(d, sink, it); }, 4, IT),
               run([&](long long* d, int it) { k_shfl<8><<<1, T>>>(d, sink, it); }, 8, IT),
               run([&](long long* d, int it) { k_mufu<1><<<1, T>>>(d, sink, it); }, 1, IT),
               run([&](long long* d, int it) { k_mufu<4><<<1, T>>>(d, sink, it); }, 4, IT),
               run([&](long long* d, int it) { k_mufu<8><<<1, T>>>(d, sink, it); }, 8, IT),
               run([&](long long* d, int it) { k_bar<<<1, T>>>(d, it); }, 1, IT));
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
