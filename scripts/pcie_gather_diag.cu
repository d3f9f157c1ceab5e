// PCIe H2D: per-piece cudaMemcpyAsync vs a zero-copy gather kernel reading pinned host memory
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

struct Piece { const uint4* src; uint4* dst; size_t n16; };
__global__ void gather_copy(const Piece* pc, int npieces, int ctas_per_piece) {
    const int p = blockIdx.x / ctas_per_piece, part = blockIdx.x % ctas_per_piece;
    if (p >= npieces) return;
    const Piece q = pc[p];
    for (size_t i = (size_t)part * blockDim.x + threadIdx.x; i < q.n16; i += (size_t)ctas_per_piece * blockDim.x) q.dst[i] = q.src[i];
}
// persistent: grid-stride over a flat list of 16-byte words (prefix table)
__global__ void gather_copy_flat(const Piece* pc, const size_t* pre, int npieces, size_t total16, int unroll) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    int p = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total16; i += stride) {
        while (pre[p + 1] <= i) p++;
        pc[p].dst[i - pre[p]] = pc[p].src[i - pre[p]];
    }
}

int main() {
    const size_t MB = 1 << 20;
    std::vector<size_t> sizes;
    for (int i = 0; i < 128; i++) { sizes.push_back((size_t)(3.07 * MB) & ~15ull); sizes.push_back((size_t)(0.64 * MB) & ~15ull); sizes.push_back((size_t)(2.56 * MB) & ~15ull); }
    const int n = (int)sizes.size();
    std::vector<void*> hs(n), ds(n);
    size_t total = 0;
    for (int i = 0; i < n; i++) { CK(cudaHostAlloc(&hs[i], sizes[i], cudaHostAllocDefault)); CK(cudaMalloc(&ds[i], sizes[i])); total += sizes[i]; }
    cudaStream_t st; CK(cudaStreamCreate(&st));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float ms;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(a, st));
        for (int i = 0; i < n; i++) CK(cudaMemcpyAsync(ds[i], hs[i], sizes[i], cudaMemcpyHostToDevice, st));
        CK(cudaEventRecord(b, st)); CK(cudaStreamSynchronize(st)); CK(cudaEventElapsedTime(&ms, a, b));
    }
    printf("cudaMemcpyAsync x%d: %.2f ms %.1f GB/s\n", n, ms, total / ms / 1e6);
    std::vector<Piece> pc(n); std::vector<size_t> pre(n + 1, 0);
    for (int i = 0; i < n; i++) {
        void* dp = nullptr; CK(cudaHostGetDevicePointer(&dp, hs[i], 0));
        pc[i] = Piece{ (const uint4*)dp, (uint4*)ds[i], sizes[i] / 16 }; pre[i + 1] = pre[i] + sizes[i] / 16;
    }
    Piece* dpc; size_t* dpre; CK(cudaMalloc(&dpc, sizeof(Piece) * n)); CK(cudaMalloc(&dpre, sizeof(size_t) * (n + 1)));
    CK(cudaMemcpy(dpc, pc.data(), sizeof(Piece) * n, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dpre, pre.data(), sizeof(size_t) * (n + 1), cudaMemcpyHostToDevice));
    for (int cpp : { 1, 2, 4 }) for (int tpb : { 256, 1024 }) {
        for (int rep = 0; rep < 3; rep++) {
            CK(cudaEventRecord(a, st));
            gather_copy<<<n * cpp, tpb, 0, st>>>(dpc, n, cpp);
            CK(cudaEventRecord(b, st)); CK(cudaStreamSynchronize(st)); CK(cudaEventElapsedTime(&ms, a, b));
        }
        printf("gather kernel, %d CTAs/piece x %d threads: %.2f ms %.1f GB/s\n", cpp, tpb, ms, total / ms / 1e6);
    }
    for (int ctas : { 8, 16, 32, 64, 148 }) {
        for (int rep = 0; rep < 3; rep++) {
            CK(cudaEventRecord(a, st));
            gather_copy_flat<<<ctas, 512, 0, st>>>(dpc, dpre, n, pre[n], 1);
            CK(cudaEventRecord(b, st)); CK(cudaStreamSynchronize(st)); CK(cudaEventElapsedTime(&ms, a, b));
        }
        printf("flat gather kernel, %d CTAs x 512: %.2f ms %.1f GB/s\n", ctas, ms, total / ms / 1e6);
    }
    return 0;
}
