// Single-GPU numerics + speed check of the tcgen05 fp8 GEMM (no communication: gating off).
// Verifies sampled outputs against a double-precision dot product of the dequantized inputs
// (the plain fp32-style reference of the same op), then times it with CUDA events.
#include <cuda_runtime.h>
#include <cuda_fp8.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../../src/kernels/gemm_fp8_sm100.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); return 2; } } while (0)

static float e4m3_to_float(uint8_t v) {
    int s = v >> 7, e = (v >> 3) & 15, m = v & 7;
    float f;
    if (e == 0) f = ldexpf((float)m, -9);            // subnormal: m * 2^-9
    else f = ldexpf(1.0f + m / 8.0f, e - 7);
    return s ? -f : f;
}

int main(int argc, char **argv) {
    int M = argc > 1 ? atoi(argv[1]) : 1024, N = argc > 2 ? atoi(argv[2]) : 1024, K = argc > 3 ? atoi(argv[3]) : 1024;
    int iters = argc > 4 ? atoi(argv[4]) : 10;
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    std::vector<uint8_t> hx((size_t)M * K), hw((size_t)N * K), hy((size_t)M * N);
    unsigned s = 12345;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return s >> 8; };
    // e4m3 codes with exponent <= 8 (|v| < 4), never NaN (0x7f / 0xff)
    for (auto &v : hx) { unsigned r = rnd(); v = (uint8_t)(((r & 1) << 7) | ((r >> 1) % 0x48)); }
    for (auto &v : hw) { unsigned r = rnd(); v = (uint8_t)(((r & 1) << 7) | ((r >> 1) % 0x48)); }
    uint8_t *dx, *dw, *dy;
    CK(cudaMalloc(&dx, hx.size())); CK(cudaMalloc(&dw, hw.size())); CK(cudaMalloc(&dy, hy.size()));
    CK(cudaMemcpy(dx, hx.data(), hx.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dw, hw.data(), hw.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dy, 0x7f, hy.size()));
    const float scale = 1.0f / sqrtf((float)K);
    acxg::Gating g = {};
    int rc = acxg::launch_gemm_fp8(dx, dw, dy, M, N, K, scale, g, sms, 0);
    if (rc) { printf("launch failed rc=%d\n", rc); return 2; }
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hy.data(), dy, hy.size(), cudaMemcpyDeviceToHost));
    int bad = 0, off1 = 0, checked = 0;
    double maxrel = 0;
    for (int t = 0; t < 4096; t++) {
        int i = rnd() % M, j = rnd() % N;
        if (t < 64) { i = (t * 37) % M; j = (t < 32 ? t : N - 1 - t) % N; }
        double acc = 0;
        for (int k = 0; k < K; k++) acc += (double)e4m3_to_float(hx[(size_t)i * K + k]) * (double)e4m3_to_float(hw[(size_t)j * K + k]);
        float ref = (float)(acc * scale);
        uint8_t want = (uint8_t)__nv_cvt_float_to_fp8(ref, __NV_SATFINITE, __NV_E4M3);
        uint8_t got = hy[(size_t)i * N + j];
        checked++;
        if (got != want) {
            float gf = e4m3_to_float(got), wf = e4m3_to_float(want);
            double rel = fabs(gf - ref) / (fabs(ref) + 1e-3);
            if (rel > maxrel) maxrel = rel;
            if (fabs(gf - wf) <= fabs(wf) * 0.13 + 2e-3) off1++;   // neighbouring e4m3 code (rounding tie / accumulation order)
            else { if (bad < 8) printf("MISMATCH (%d,%d): got 0x%02x (%g) want 0x%02x (%g) ref %g\n", i, j, got, gf, want, wf, ref); bad++; }
        }
    }
    printf("gemm_fp8 %dx%dx%d: checked %d, exact %d, neighbour-code %d, wrong %d\n", M, N, K, checked, checked - off1 - bad, off1, bad);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; i++) acxg::launch_gemm_fp8(dx, dw, dy, M, N, K, scale, g, sms, 0);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; i++) acxg::launch_gemm_fp8(dx, dw, dy, M, N, K, scale, g, sms, 0);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double tf = 2.0 * M * N * K / (ms / iters * 1e-3) / 1e12;
    printf("{\"bench\": \"gemm_fp8\", \"M\": %d, \"N\": %d, \"K\": %d, \"ms\": %.4f, \"tflops\": %.1f, \"wrong\": %d}\n", M, N, K, ms / iters, tf, bad);
    printf("gemm_fp8_test: %s\n", bad == 0 && off1 < checked / 20 ? "PASS" : "FAIL");
    return !(bad == 0 && off1 < checked / 20);
}
