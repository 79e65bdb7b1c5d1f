// Layout microbenchmark ("decide the layout on ncu evidence").
// Runs the Miller loop's dominant operation mix (Fp12 squaring + sparse line products) for many iterations per thread in
//   A. the round-1 layout: one proof per thread, tower operands behind pointers into the per-thread stack, 3 blocks of 128 per SM;
//   B. the shared-memory-resident lazily reduced layout of csrc/lazy.cuh: 2 blocks of 128 per SM, 28 slots of shared memory per thread;
// checks that both produce identical values, and prints time, executed IMAD.WIDE per thread and the fraction of the IMAD.WIDE issue
// rate (measured in the same run by k_imad_wide) each one reaches.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/build/lzbench tools/lzbench.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "../stylus_zkvm_verifiers_b200/csrc/kernels.cuh"
#include "../stylus_zkvm_verifiers_b200/csrc/lazy.cuh"
#include "../stylus_zkvm_verifiers_b200/csrc/lazy2.cuh"
using namespace zkv;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

// mode 0: f <- f^2 ; mode 1: f <- f^2 * (1 + (c3 + c4 v) w) twice (the fixed-pair part of a Miller doubling step)
__global__ void __launch_bounds__(128, 3) k_old(int iters, int mode, const fp12* in, const fp2* cs, fp12* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    fp12 f = in[i]; fp2 c3 = cs[2 * i], c4 = cs[2 * i + 1];
    for (int k = 0; k < iters; k++) {
        f12_sqr(f, f);
        if (mode) { f12_mul_line1(f, c3, c4); f12_mul_line1(f, c4, c3); }
    }
    out[i] = f;
}
// Phase skew experiment: every second block that lands on an SM spins for `g_skew` cycles first, so that the two warps of a scheduler
// (one per block) are not in the same phase of the same instruction stream (multiplying together, then adding together).
__device__ int g_skew = 0;
__device__ int g_smcount[256];
__device__ __forceinline__ void phase_skew() {
    if (g_skew == 0) return;
    __shared__ int role;
    if (threadIdx.x == 0) { unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); role = atomicAdd(&g_smcount[sm & 255], 1) & 1; }
    __syncthreads();
    if (role) { long long t0 = clock64(); while (clock64() - t0 < g_skew) { } }
}
template <int MINB>
__global__ void __launch_bounds__(LZ_NT, MINB) k_lz(int iters, int mode, const fp12* in, const fp2* cs, fp12* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    phase_skew();
    const uint32_t F = lz_tid(), T = F + 12 * LZ_SLOT, L = F + 24 * LZ_SLOT, L2 = F + 18 * LZ_SLOT;
    { fp12 f = in[i]; const fp* w = &f.c0.c0.c0; for (int k = 0; k < 12; k++) lz_stfp(F + k * LZ_SLOT, w[k]); }
    lz_st2(L, cs[2 * i]); lz_st2(L + 2 * LZ_SLOT, cs[2 * i + 1]);
    lz_st2(L2, cs[2 * i + 1]); lz_st2(L2 + 2 * LZ_SLOT, cs[2 * i]);
    for (int k = 0; k < iters; k++) {
        lz_f12sqr(F, T);
        if (mode) { lz_mul_nline(F, T, L); lz_mul_nline(F, T, L2); }
    }
    fp12 f; fp* w = &f.c0.c0.c0; for (int k = 0; k < 12; k++) w[k] = lz_ldfp(F + k * LZ_SLOT);
    out[i] = f;
}

//   C. two lanes per proof (csrc/lazy2.cuh): 2 blocks of 256 threads = 128 proofs per SM-block, the same 28 slots per proof
__global__ void __launch_bounds__(2 * LZ_NT, 2) k_lz2(int iters, int mode, const fp12* in, const fp2* cs, fp12* out) {
    const Lz2 c = lz2_ctx();
    const int i = blockIdx.x * LZ_NT + lz2_pid();
    const uint32_t F = lz2_pid(), T = F + 12 * LZ_SLOT, L = F + 24 * LZ_SLOT, L2 = F + 18 * LZ_SLOT;
    { const fp* w = &in[i].c0.c0.c0; for (int k = 0; k < 6; k++) lz2_sto(F, k, c, w[2 * k + c.im]); }
    { const fp* a = &cs[2 * i].c0; const fp* b = &cs[2 * i + 1].c0; lz2_sto(L, 0, c, a[c.im]); lz2_sto(L, 1, c, b[c.im]); lz2_sto(L2, 0, c, b[c.im]); lz2_sto(L2, 1, c, a[c.im]); }
    lz2_sync();
    for (int k = 0; k < iters; k++) {
        lz2_f12sqr(F, T, c);
        if (mode) { lz2_mul_nline(F, T, L, c); lz2_mul_nline(F, T, L2, c); }
    }
    fp* w = &out[i].c0.c0.c0; for (int k = 0; k < 6; k++) w[2 * k + c.im] = lz2_ldo(F, k, c);
}

static uint64_t sm64(uint64_t& s) { uint64_t z = (s += 0x9e3779b97f4a7c15ull); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); }

int main(int argc, char** argv) {
    int iters = argc > 1 ? atoi(argv[1]) : 64;
    int skew = argc > 2 ? atoi(argv[2]) : 0;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const int n = sms * 3 * 128 * 2;     // enough threads for two full waves of either layout's largest grid
    std::vector<fp12> h_in(n); std::vector<fp2> h_cs(2 * n);
    uint64_t seed = 0xB200;
    auto rnd_fp = [&](fp& x) { for (int k = 0; k < 8; k++) x.v[k] = (uint32_t)sm64(seed); x.v[7] &= 0x1fffffffu; };   // < 2^253 < p: a canonical residue
    for (auto& f : h_in) { fp* w = &f.c0.c0.c0; for (int k = 0; k < 12; k++) rnd_fp(w[k]); }
    for (auto& c : h_cs) { rnd_fp(c.c0); rnd_fp(c.c1); }
    fp12 *d_in, *d_o1, *d_o2, *d_o3; fp2* d_cs;
    CK(cudaMalloc(&d_in, n * sizeof(fp12))); CK(cudaMalloc(&d_o1, n * sizeof(fp12))); CK(cudaMalloc(&d_o2, n * sizeof(fp12))); CK(cudaMalloc(&d_o3, n * sizeof(fp12))); CK(cudaMalloc(&d_cs, 2 * n * sizeof(fp2)));
    CK(cudaMemcpy(d_in, h_in.data(), n * sizeof(fp12), cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_cs, h_cs.data(), 2 * n * sizeof(fp2), cudaMemcpyHostToDevice));
    CK(cudaMemcpyToSymbol(g_skew, &skew, sizeof skew));
    const size_t smem = (size_t)LZ_SLOTS * 32 * LZ_NT;
    const int lzb = 256 / LZ_NT;        // blocks per SM of the shared-memory layout (256 threads per SM either way)
    CK(cudaFuncSetAttribute(k_lz<256 / LZ_NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_lz<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_lz2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ_lz2 = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_lz2, k_lz2, 2 * LZ_NT, smem));
    int occ_old = 0, occ_lz = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_old, k_old, 128, 0));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_lz, k_lz<256 / LZ_NT>, LZ_NT, smem)); (void)lzb;
    // IMAD.WIDE issue roofline, measured here
    uint32_t* d_w; CK(cudaMalloc(&d_w, (size_t)sms * 8 * 256 * 4));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < 4; r++) { CK(cudaEventRecord(e0)); k_imad_wide<<<sms * 8, 256>>>(d_w, 0x9e3779b9u, 0x7f4a7c15u, 4096); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (r) best = ms < best ? ms : best; }
    const double peak = (double)sms * 8 * 256 * 4096 * 64.0 / (best * 1e-3);
    printf("{\"sms\": %d, \"imad_wide_peak_per_s\": %.4g, \"blocks_per_sm\": {\"old\": %d, \"lz\": %d}, \"iters\": %d, \"phase_skew_cycles\": %d, \"runs\": [\n", sms, peak, occ_old, occ_lz, iters, skew);
    // IMAD.WIDE per iteration (static counts: Fp2 product of the round-1 tower 336 = 3 x 64 + 2 x 72; lazily reduced Fp6 product 1584, sparse 01-product 1392)
    const double mac_old[2] = {12 * 336.0, 12 * 336.0 + 2 * 10 * 336.0}, mac_lz[2] = {2 * 1584.0, 2 * 1584.0 + 4 * 1392.0};
    const double mac_lz2[2] = {2 * 2 * 984.0, 2 * 2 * 984.0 + 4 * 2 * 856.0};      // per PROOF (both lanes): dense 2 x (6 x 128 + 3 x 72), sparse 2 x (5 x 128 + 3 x 72)
    bool first = true; int bad = 0;
    for (int mode = 0; mode < 2; mode++) {
        for (int variant = 0; variant < 4; variant++) {
            int per_sm = variant == 0 ? occ_old : variant == 1 ? occ_lz : variant == 3 ? occ_lz2 : 1;
            int blocks = sms * per_sm, threads = blocks * (variant == 0 ? 128 : LZ_NT);      // "threads" = proofs in flight
            float tb = 1e30f;
            for (int r = 0; r < 3; r++) {
                CK(cudaEventRecord(e0));
                if (variant == 0) k_old<<<blocks, 128>>>(iters, mode, d_in, d_cs, d_o1);
                else if (variant == 1) k_lz<256 / LZ_NT><<<blocks, LZ_NT, smem>>>(iters, mode, d_in, d_cs, d_o2);
                else if (variant == 3) k_lz2<<<blocks, 2 * LZ_NT, smem>>>(iters, mode, d_in, d_cs, d_o3);
                else k_lz<1><<<blocks, LZ_NT, smem>>>(iters, mode, d_in, d_cs, d_o2);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (r) tb = ms < tb ? ms : tb;
            }
            double mac = (variant == 0 ? mac_old[mode] : variant == 3 ? mac_lz2[mode] : mac_lz[mode]) * iters * threads;
            printf("%s {\"mode\": \"%s\", \"layout\": \"%s\", \"blocks_per_sm\": %d, \"threads\": %d, \"ms\": %.4f, \"iters_per_s\": %.4g, \"imad_wide_per_thread_iter\": %.0f, \"executed_frac_of_imad_peak\": %.4f}",
                   first ? "" : ",\n", mode ? "sqr+2nline" : "sqr", variant == 0 ? "r1_thread_stack" : variant == 3 ? "smem_lazy_2lanes" : "smem_lazy", per_sm, threads, tb, (double)threads * iters / (tb * 1e-3),
                   variant == 0 ? mac_old[mode] : variant == 3 ? mac_lz2[mode] : mac_lz[mode], mac / (tb * 1e-3) / peak);
            first = false;
            if (variant == 1) {      // same inputs, same iteration count: the two layouts must agree bit for bit (compare the common prefix of threads)
                int m = sms * (occ_old < occ_lz ? occ_old : occ_lz) * 128;
                std::vector<fp12> a(m), b(m);
                CK(cudaMemcpy(a.data(), d_o1, m * sizeof(fp12), cudaMemcpyDeviceToHost)); CK(cudaMemcpy(b.data(), d_o2, m * sizeof(fp12), cudaMemcpyDeviceToHost));
                if (memcmp(a.data(), b.data(), m * sizeof(fp12)) != 0) bad++;
            }
            if (variant == 3) {      // the two-lane layout against the one-thread shared-memory layout (run just before, same iteration count)
                int m = sms * (occ_lz2 < occ_lz ? occ_lz2 : occ_lz) * 128;
                std::vector<fp12> a(m), b(m);
                CK(cudaMemcpy(a.data(), d_o2, m * sizeof(fp12), cudaMemcpyDeviceToHost)); CK(cudaMemcpy(b.data(), d_o3, m * sizeof(fp12), cudaMemcpyDeviceToHost));
                if (memcmp(a.data(), b.data(), m * sizeof(fp12)) != 0) { bad++; fprintf(stderr, "two-lane layout differs (mode %d)\n", mode); }
            }
        }
    }
    printf("\n], \"layouts_agree\": %s}\n", bad ? "false" : "true");
    return bad ? 2 : 0;
}
