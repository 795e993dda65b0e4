// Cost of an (almost) empty level of rbs_ent_levels: what is the floor per level?
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "../../ribasim_b200/csrc/rbs_kernels.cuh"
__global__ void __launch_bounds__(1024) k_lv(EntSched es, const uint16_t* blob, long long* clk, int variant) {
    extern __shared__ double smem[];
    double* S = smem;
    uint16_t* ix = (uint16_t*)(S + (size_t)es.n_slots * 8);
    for (int w = threadIdx.x; w < es.total; w += blockDim.x) ix[w] = blob[w];
    for (int w = threadIdx.x; w < es.n_slots * 8; w += blockDim.x) S[w] = 1.0 + w * 1e-6;
    __syncthreads();
    long long t0 = clock64();
    if (variant == 0) rbs_ent_levels(S, ix, es, 0, es.n_elim, 3u);
    else if (variant == 1) { for (int lv = 0; lv < es.n_elim; lv++) __syncthreads(); }
    else if (variant == 2) {
        const uint16_t *lvl = ix + es.off_lvl, *team = ix + es.off_team;
        int acc = 0;
        for (int lv = 0; lv < es.n_elim; lv++) { acc += lvl[lv] + lvl[lv + 1] + team[lv]; __syncthreads(); }
        if (acc == 12345) S[0] = 1;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) clk[0] = t1 - t0;
}
int main() {
    const int NL = 64;
    for (int ne : {0, 1, 8, 64, 256}) {
        // NL levels with `ne` entries each, one product per entry, team 0
        std::vector<uint16_t> b;
        EntSched es{};
        int n_ent = ne * NL;
        for (int k = 0; k < n_ent; k++) b.push_back((uint16_t)(100 + k % 1000));           // ent_e
        es.off_q0 = b.size(); for (int k = 0; k < n_ent; k++) b.push_back((uint16_t)k);    // q0
        es.off_cf = b.size(); for (int k = 0; k < n_ent; k++) b.push_back(1);              // 1 product
        es.off_pl = b.size(); for (int k = 0; k < n_ent; k++) b.push_back((uint16_t)(k % 97));
        es.off_pp = b.size(); for (int k = 0; k < n_ent; k++) b.push_back((uint16_t)(k % 89));
        es.off_pu = b.size(); for (int k = 0; k < n_ent; k++) b.push_back((uint16_t)(k % 83));
        es.off_lvl = b.size(); for (int l = 0; l <= NL; l++) b.push_back((uint16_t)(l * ne));
        es.off_team = b.size(); for (int l = 0; l < NL; l++) b.push_back(0);
        es.off_perm = b.size();
        while (b.size() % 8) b.push_back(0);
        es.total = b.size(); es.n_elim = NL; es.n_bwd = 0; es.n_slots = 2878; es.y0 = 2366; es.dbg = 0;
        uint16_t* d; long long* clk;
        cudaMalloc(&d, b.size() * 2); cudaMemcpy(d, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
        cudaMalloc(&clk, 8);
        size_t sm = (size_t)es.n_slots * 64 + b.size() * 2;
        cudaFuncSetAttribute(k_lv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        for (int variant : {0, 1, 2}) {
            for (int rep = 0; rep < 2; rep++) k_lv<<<1, 1024, sm>>>(es, d, clk, variant);
            cudaDeviceSynchronize();
            long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
            printf("entries/level=%d variant=%d: %.1f cycles per level  (%s)\n", ne, variant, h / (double)NL, cudaGetErrorString(cudaGetLastError()));
        }
        cudaFree(d); cudaFree(clk);
    }
    return 0;
}
