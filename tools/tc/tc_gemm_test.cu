// Standalone check of blockdft_tc_kernel (tcgen05 3xTF32 GEMM) against a float64 CPU product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tc/tc_gemm_test tools/tc/tc_gemm_test.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include "../../audio-analyzer-omega_b200/csrc/blockdft_tc_kernel.cuh"
using namespace o4;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

int main() {
    const int hop = 512, nb = 1000, n_ch = 3, cols = 512, halves = cols / TC_BN;
    const int j0 = 0;
    std::vector<float> x((size_t)n_ch * nb * hop), E((size_t)hop * cols);
    srand(1);
    for (auto& v : x) v = (float)rand() / RAND_MAX - 0.5f;
    for (int k = 0; k < hop; ++k) for (int c = 0; c < cols; ++c) E[(size_t)k * cols + c] = (float)cos(0.001 * (k + 1) * (c + 3) + c);
    const int nkc = hop / TC_KC;
    std::vector<uint8_t> img((size_t)halves * nkc * 2 * TC_B_BYTES, 0);
    for (int h = 0; h < halves; ++h) for (int kc = 0; kc < nkc; ++kc) for (int n = 0; n < TC_BN; ++n) for (int e = 0; e < TC_KC; ++e) {
        const float v = E[(size_t)(kc * TC_KC + e) * cols + h * TC_BN + n];
        uint32_t u; memcpy(&u, &v, 4); u &= 0xFFFFE000u; float hi; memcpy(&hi, &u, 4);
        const float lo = v - hi;
        const size_t base = ((size_t)(h * nkc + kc) * 2) * TC_B_BYTES;
        const int off = tc_sw64_offset(n, e >> 2) + (e & 3) * 4;
        memcpy(&img[base + off], &hi, 4);
        memcpy(&img[base + TC_B_BYTES + off], &lo, 4);
    }
    float *dx, *dq; uint8_t* dimg;
    CK(cudaMalloc(&dx, x.size() * 4)); CK(cudaMalloc(&dq, (size_t)n_ch * nb * cols * 4)); CK(cudaMalloc(&dimg, img.size()));
    CK(cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dimg, img.data(), img.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dq, 0xff, (size_t)n_ch * nb * cols * 4));
    BlockDftTcArgs a; memset(&a, 0, sizeof a);
    a.x = dx; a.ch_stride = (long long)nb * hop; a.hop = hop; a.n_ch = n_ch; a.j0 = j0; a.nb = nb; a.n_halves = halves;
    a.Eimg = dimg; a.Q = dq; a.qs = cols;
    const size_t smem = blockdft_tc_smem_bytes();
    CK(cudaFuncSetAttribute(blockdft_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = ((nb + TC_BM - 1) / TC_BM) * n_ch * halves;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    blockdft_tc_kernel<<<grid, TC_THREADS + 32, smem>>>(a);
    CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) blockdft_tc_kernel<<<grid, TC_THREADS + 32, smem>>>(a);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<float> q((size_t)n_ch * nb * cols);
    CK(cudaMemcpy(q.data(), dq, q.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0; long bad = 0;
    for (int c = 0; c < n_ch; ++c) for (int r = 0; r < nb; r += 7) for (int n = 0; n < cols; n += 5) {
        double acc = 0;
        for (int k = 0; k < hop; ++k) acc += (double)x[((size_t)c * nb + r) * hop + k] * (double)E[(size_t)k * cols + n];
        const double got = q[((size_t)c * nb + r) * cols + n];
        const double err = fabs(got - acc);
        if (!(err < 1e-3)) { if (bad < 10) printf("bad c=%d r=%d n=%d got=%g ref=%g\n", c, r, n, got, acc); ++bad; }
        if (err > maxerr) maxerr = err;
        if (fabs(acc) > maxref) maxref = fabs(acc);
    }
    printf("tc gemm: max abs err %.3e (max |ref| %.3f), bad %ld, %.3f ms per launch of %d x %d x %d\n", maxerr, maxref, bad, ms / 10, n_ch * nb, cols, hop);
    return bad ? 2 : 0;
}
