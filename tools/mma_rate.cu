// tools/mma_rate.cu -- what the legacy tensor path (mma.sync.m16n8k16 bf16, SASS HMMA) sustains on this GPU, operands in registers and with
// the B fragment re-read from shared memory before every MMA (2 x LDS.32 per lane).  Decides whether a hand-written mma.sync kernel can
// beat cuBLASLt + a separate elementwise kernel for the policy's gated FFN.   nvcc -arch=sm_100a -O3 -o /tmp/mma_rate tools/mma_rate.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int SMEM_B, int MT>
__global__ void __launch_bounds__(256) k_rate(float* out, int iters) {
    __shared__ uint32_t bs[64 * 36];
    for (int i = threadIdx.x; i < 64 * 36; i += 256) bs[i] = 0x3f803f80u;
    __syncthreads();
    const int lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
    uint32_t a[MT][4];
    float d[MT][8][4];
    for (int m = 0; m < MT; ++m) {
        for (int i = 0; i < 4; ++i) a[m][i] = 0x3f803f80u + lane + m;
        for (int j = 0; j < 8; ++j) for (int i = 0; i < 4; ++i) d[m][j][i] = 0.f;
    }
    uint32_t b0 = 0x3f803f80u, b1 = 0x3f803f80u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (SMEM_B) { b0 = bs[(j * 8 + g) * 36 + tg + (it & 3) * 8]; b1 = bs[(j * 8 + g) * 36 + tg + 4 + (it & 3) * 8]; }
#pragma unroll
            for (int m = 0; m < MT; ++m) mma(d[m][j], a[m], b0, b1);
        }
    }
    float acc = 0.f;
    for (int m = 0; m < MT; ++m) for (int j = 0; j < 8; ++j) for (int i = 0; i < 4; ++i) acc += d[m][j][i];
    if (acc == 12345.f) out[0] = acc;
}

template <int SMEM_B, int MT>
void run(const char* what, int sms) {
    float* out; cudaMalloc(&out, 4);
    const int iters = 4096, blocks = sms * 2;
    k_rate<SMEM_B, MT><<<blocks, 256>>>(out, 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_rate<SMEM_B, MT><<<blocks, 256>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flop = (double)blocks * 8 * iters * 8 * MT * 4096.0;
    printf("%-60s %8.1f TFLOP/s  (%.3f ms, %s)\n", what, flop / (ms * 1e-3) / 1e12, ms, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("mma.sync.m16n8k16 bf16 -> fp32, %d SMs, 2 blocks x 8 warps per SM\n", sms);
    run<0, 1>("operands in registers, 8 independent accumulators / warp", sms);
    run<0, 2>("operands in registers, 16 independent accumulators / warp", sms);
    run<1, 1>("B fragment from shared memory per MMA (1 m-tile per B)", sms);
    run<1, 2>("B fragment from shared memory per 2 MMAs (2 m-tiles per B)", sms);
    return 0;
}
